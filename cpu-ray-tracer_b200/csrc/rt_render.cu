// rt_render.cu — the two integrators as wavefront pipelines over SoA ray queues.
//
// Replaces (reference, /root/reference):
//   path tracer  3. PathTracer/renderer.cpp: Sample :50-100, HandleMirror :20-25, HandleDielectric :27-45,
//                ProcessTile :117-131, Tick :144-168; diffusereflection template/tmplmath.h:535-544
//   Whitted      2. WhittedStyle/renderer.cpp: Trace :21-91, DirectIllumination :105-126, Tick :131-157
//
// Path tracer.  The reference carries ONE xorshift32 stream through the 256 pixels of a 16x16 tile
// (seeded per tile per frame), so pixels of a tile are serially dependent but every (tile, frame) pair
// is independent.  Each such pair is one wavefront *slot*: it owns one ray at a time; when its path
// ends the shade stage splats the sample and regenerates the next pixel's primary ray from the same
// stream.  Stages (separate kernels, persistent grid-stride, counts stay on the device):
//   generate  first primary ray of every slot
//   extend    closest hit for every queued slot (rt_device.cuh traverse)
//   shade     sky / light / depth-limit termination, material evaluation, RNG, next ray, regeneration;
//             survivors are compacted into the next queue with warp ballot/popc + one atomic per warp
// The reference path tracer has no next-event estimation, so it has no connect stage.
//
// Whitted.  Rays carry (pixel, weight); shade splats sky / light / ambient terms, pushes reflection and
// refraction rays into the next queue and one shadow ray per diffuse hit into the shadow queue;
//   connect   any-hit occlusion kernel over the shadow queue, adds the direct term when visible.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rt_internal.h"

namespace rtb {

struct PTState {
    float4* rayO;      // O.xyz, -
    float4* rayD;      // D.xyz, int flags = inside | depth << 8
    float4* hit;       // t, u, v, int objIdx
    int* hitTri;
    uint32_t* seed;
    int* pix;          // next SAMPLE of the tile this slot will generate (0..256 * passes); its pixel = sample / passes
    float4* weights;   // [depth_limit][slots]: throughput factor of every bounce (see k_pt_shade)
    int* active[2];
    int* count;        // [2]
    unsigned long long* counters; // [0] extension rays, [1] shadow rays, [2] iterations
    int* history;      // queued rays per iteration of the current batch
    int iteration;
    float4* accum;
    float4* frameBuf;  // optional: one float4 image per sample of the launch (ordered accumulation / look-ahead mode), else nullptr
    // layout of an image in frameBuf: 0 = W x H pixels; 1 = compact, only the job's own tiles: (k-th tile of the job) * 256 + pixel of the
    // tile (k_pt_streams8 + k_sum_frames: a tile shard of N ranks needs 1 / N of the memory, and a tile's samples are contiguous)
    int frameCompact;
    float invTileStep; // 1 / tileStep (exact tile -> k for the job's tiles: the quotient is an integer below 2^22)
    int slots, nTiles, tilesX, tileBegin, tileStep; // the job's k-th tile is tileBegin + k * tileStep
    int firstSpp, stride;
    int W, H, depthLimit, seedMode;
    int passes;        // Renderer::passes: samples per pixel per frame, consecutive in the tile's stream (renderer.cpp:123)
    float eps;
};

__device__ __forceinline__ uint32_t pt_seed(const PTState& p, int tile, int spp)
{
    const int tx = tile % p.tilesX, ty = tile / p.tilesX;
    return init_seed((uint32_t)(tx + ty * p.W + spp * 1799)); // renderer.cpp:120
}

// Generates the primary ray of pixel `pix` of the slot's tile (renderer.cpp:121-126): jitter draws
// y first, then x (argument evaluation order of the reference build, see oracle/ref_build).
__device__ __forceinline__ uint32_t pt_pixel_seed(const PTState& p, int x, int y, int spp)
{
    uint32_t seed = init_seed((uint32_t)(x + y * p.W) + (uint32_t)spp * 0x9E3779B1u);
    return seed == 0 ? 0x12345678u : seed;
}

// Wavefront slots.  RT_SEED_REFERENCE_TILE: slot = (tile, frame), it walks the tile's 256 pixels with the tile's stream.
// RT_SEED_PER_PIXEL: slot = (tile, frame, pixel) with the pixel's own stream - every pixel of every frame in flight is
// its own slot, so a batch needs only passes x (depthLimit + 1) extend / shade iterations.
__device__ __forceinline__ void pt_slot(const PTState& p, int slot, int& tile, int& frame, int& px0)
{
    const bool perPixel = p.seedMode == RT_SEED_PER_PIXEL;
    const int unit = perPixel ? slot >> 8 : slot;
    px0 = perPixel ? slot & 255 : 0;
    tile = p.tileBegin + (unit % p.nTiles) * p.tileStep;
    frame = unit / p.nTiles;
}

__device__ __forceinline__ void pt_generate(const PTState& p, const DCamera& cam, int slot, int pix, uint32_t& seed, float3& D)
{
    int tile, frame, px0;
    pt_slot(p, slot, tile, frame, px0);
    const int tx = tile % p.tilesX, ty = tile / p.tilesX;
    const int x = tx * 16 + (pix & 15), y = ty * 16 + (pix >> 4);
    const float jy = random_float(seed);
    const float jx = random_float(seed);
    D = primary_dir(cam, (float)x + jx, (float)y + jy);
}

__global__ void __launch_bounds__(256) k_pt_generate(const PTState p, const DCamera cam)
{
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < p.slots; slot += gridDim.x * blockDim.x)
    {
        int tile, frame, px0;
        pt_slot(p, slot, tile, frame, px0);
        const int spp = p.firstSpp + frame * p.stride;
        const int tx = tile % p.tilesX, ty = tile / p.tilesX;
        uint32_t seed = p.seedMode == RT_SEED_PER_PIXEL ? pt_pixel_seed(p, tx * 16 + (px0 & 15), ty * 16 + (px0 >> 4), spp) : pt_seed(p, tile, spp);
        float3 D;
        pt_generate(p, cam, slot, px0, seed, D);
        p.rayO[slot] = make_float4(cam.pos.x, cam.pos.y, cam.pos.z, 0);
        p.rayD[slot] = make_float4(D.x, D.y, D.z, __int_as_float(0));
        p.seed[slot] = seed;
        p.pix[slot] = px0 * p.passes + 1; // index of the next sample
        p.active[0][slot] = slot;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) p.count[0] = p.slots, p.count[1] = 0, p.count[4] = 0;
}

template <int ACCEL>
__global__ void __launch_bounds__(128) k_pt_extend(const PTState p, const DScene s, int cur)
{
    const int n = p.count[cur];
    const int* __restrict__ active = p.active[cur];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int slot = active[i];
        const float4 o = p.rayO[slot], d = p.rayD[slot];
        HitRec h;
        find_nearest<false, ACCEL>(s, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), 1e34f, h);
        p.hit[slot] = make_float4(h.t, h.u, h.v, __int_as_float(h.obj));
        p.hitTri[slot] = h.tri;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        p.count[cur ^ 1] = 0; // the shade stage of this iteration appends here
        p.counters[0] += (unsigned long long)n;
        p.counters[2] += 1;
        p.history[p.iteration] = n;
    }
}

// extend, persistent-warp version (rt_device.cuh trace_queue): lanes pull rays from the compacted
// slot queue as they free up
struct PTSrc {
    const PTState& p;
    const int* __restrict__ active;
    __device__ __forceinline__ bool load(int i, float3& O, float3& D, float& tmax) const
    {
        const int slot = active[i];
        const float4 o = p.rayO[slot], d = p.rayD[slot];
        O = f3(o.x, o.y, o.z), D = f3(d.x, d.y, d.z), tmax = 1e34f;
        return true;
    }
    __device__ __forceinline__ void world(int i, float3& O, float3& D) const
    {
        const int slot = active[i];
        const float4 o = p.rayO[slot], d = p.rayD[slot];
        O = f3(o.x, o.y, o.z), D = f3(d.x, d.y, d.z);
    }
    __device__ __forceinline__ void store(int i, const HitRec& h) const
    {
        const int slot = active[i];
        p.hit[slot] = make_float4(h.t, h.u, h.v, __int_as_float(h.obj));
        p.hitTri[slot] = h.tri;
    }
};

template <int ACCEL>
__global__ void __launch_bounds__(128) k_pt_extend_persistent(const PTState p, const DScene s, int cur)
{
    const int n = p.count[cur];
    PTSrc src = { p, p.active[cur] };
    accel_trace_queue<ACCEL, false, false>(s, src, n, p.count + 4);
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        p.count[cur ^ 1] = 0;
        p.counters[0] += (unsigned long long)n;
        p.counters[2] += 1;
        p.history[p.iteration] = n;
    }
}

// diffusereflection (tmplmath.h:535-544): uniform direction by rejection; the three draws fill z, y, x
// (right-to-left argument evaluation of make_float3 in the reference build).
__device__ __forceinline__ float3 diffuse_reflection(float3 N, uint32_t& seed)
{
    float3 R;
    do
    {
        const float rz = random_float(seed) * 2 - 1;
        const float ry = random_float(seed) * 2 - 1;
        const float rx = random_float(seed) * 2 - 1;
        R = f3(rx, ry, rz);
    } while (dot(R, R) > 1);
    if (dot(R, N) < 0) R = R * -1.0f;
    return normalize(R);
}

// One bounce of Renderer::Sample (renderer.cpp:50-100) on register state, in two pieces so that the
// stream kernel can run the rejection loop of the diffuse lobe as its own warp-voted state.
//   pt_surface: depth limit :55, GetHitInfo :57-66, light :69, absorption :76-80, lobe pick :82-99.
//     PT_END   the path ends here with leaf radiance L
//     PT_NEXT  mirror / dielectric: throughput factor w and direction nD are final
//     PT_DIFF  diffuse lobe: N is the shading normal and w holds medium_scale * brdf * 2 * PI, still to
//              be multiplied by dot(nD, N) once nD is drawn (same left-to-right product as :98)
// Shared by the wavefront shade stage and the stream kernels so all evaluate the very same expressions.
enum { PT_END = 0, PT_NEXT = 1, PT_DIFF = 2 };

__device__ __forceinline__ int pt_surface(const DScene& s, const int depthLimit,
    const float3 O, const float3 D, const bool inside, const int depth,
    const float t, const float bu, const float bv, const int obj, const int tri, uint32_t& seed,
    float3& L, float3& w, float3& I, float3& N, float3& nD, bool& nInside)
{
    L = f3(0, 0, 0);
    if (depth >= depthLimit) return PT_END;
    I = O + t * D;
    ShadeHit h;
    float uu, vv;
    hit_info(s, D, I, obj, tri, bu, bv, h, uu, vv);
    if (h.isLight) { L = f3(s.light_color[0], s.light_color[1], s.light_color[2]); return PT_END; }
    float3 medium_scale = f3(1, 1, 1);
    if (inside)
    {
        const float3 a = h.absorption * -t; // renderer.cpp:76-80
        medium_scale = beer_scale(a.x, a.y, a.z);
    }
    const float r = random_float(seed);
    nInside = false;
    if (r < h.reflectivity)
    {
        nD = reflect(D, h.N); // HandleMirror :20-25
        w = h.albedo * medium_scale;
        return PT_NEXT;
    }
    if (r < h.reflectivity + h.refractivity)
    {
        // HandleDielectric :27-45
        w = h.albedo * medium_scale;
        nD = reflect(D, h.N);
        const float n1 = inside ? 1.2f : 1, n2 = inside ? 1 : 1.2f;
        const float eta = n1 / n2, cosi = dot(-D, h.N);
        const float cost2 = 1.0f - eta * eta * (1 - cosi * cosi);
        if (cost2 > 0)
        {
            const float a = n1 - n2, b = n1 + n2, R0 = (a * a) / (b * b), c = 1 - cosi;
            const float Fr = R0 + (1 - R0) * (c * c * c * c * c);
            const float3 T = eta * D + ((eta * cosi - sqrtf(fabsf(cost2))) * h.N);
            if (random_float(seed) > Fr) nD = T, nInside = !inside;
        }
        return PT_NEXT;
    }
    N = h.N;
    const float3 brdf = h.albedo * RT_INVPI;
    w = medium_scale * brdf * 2.0f * RT_PI; // renderer.cpp:98, all but the last factor
    return PT_DIFF;
}

// Returns true when the path ends here with leaf radiance L (sky :54, depth limit :55, light :69);
// otherwise the throughput factor w of this bounce and the continuation ray.
__device__ __forceinline__ bool pt_bounce(const DScene& s, const float eps, const int depthLimit,
    const float3 O, const float3 D, const bool inside, const int depth,
    const float t, const float bu, const float bv, const int obj, const int tri, uint32_t& seed,
    float3& L, float3& w, float3& nO, float3& nD, bool& nInside)
{
    if (obj == -1) { L = sky_color(s, D); return true; }
    float3 I, N;
    const int k = pt_surface(s, depthLimit, O, D, inside, depth, t, bu, bv, obj, tri, seed, L, w, I, N, nD, nInside);
    if (k == PT_END) return true;
    if (k == PT_DIFF)
    {
        nD = diffuse_reflection(N, seed);
        w = w * dot(nD, N);
    }
    nO = I + nD * eps;
    return false;
}

// shade stage of the wavefront.  The reference recursion returns w0 * (w1 * (... * L)); the factors
// w_d are stored per depth and multiplied back in that order when the path ends, so a sample is
// bit-identical to the recursive evaluation.
__global__ void __launch_bounds__(128) k_pt_shade(const PTState p, const DScene s, const DCamera cam, int cur)
{
    const int n = p.count[cur];
    const int* __restrict__ active = p.active[cur];
    int* __restrict__ nextActive = p.active[cur ^ 1];
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.count[4] = 0; // fetch counter of the next extend launch
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x)
    {
        const int i = base + threadIdx.x;
        bool alive = false;
        int slot = -1;
        if (i < n)
        {
            slot = active[i];
            alive = true;
            const float4 o4 = p.rayO[slot], d4 = p.rayD[slot], h4 = p.hit[slot];
            const int flags = __float_as_int(d4.w);
            int depth = flags >> 8;
            uint32_t seed = p.seed[slot];
            float3 L, w, nO, nD;
            bool nInside;
            const bool terminated = pt_bounce(s, p.eps, p.depthLimit, f3(o4.x, o4.y, o4.z), f3(d4.x, d4.y, d4.z), flags & 1, depth,
                h4.x, h4.y, h4.z, __float_as_int(h4.w), p.hitTri[slot], seed, L, w, nO, nD, nInside);
            if (!terminated)
            {
                p.weights[(size_t)depth * p.slots + slot] = make_float4(w.x, w.y, w.z, 0);
                depth++;
                p.rayO[slot] = make_float4(nO.x, nO.y, nO.z, 0);
                p.rayD[slot] = make_float4(nD.x, nD.y, nD.z, __int_as_float((nInside ? 1 : 0) | (depth << 8)));
            }
            else
            {
                for (int d = depth - 1; d >= 0; d--)
                {
                    const float4 wd = p.weights[(size_t)d * p.slots + slot];
                    L = f3(wd.x, wd.y, wd.z) * L;
                }
                const int smp = p.pix[slot]; // index of the NEXT sample; the finished one is smp - 1
                const int pix = (smp - 1) / p.passes + 1; // pixel after the finished sample's pixel
                int tile, frame, px0;
                pt_slot(p, slot, tile, frame, px0);
                {
                    const int tx = tile % p.tilesX, ty = tile / p.tilesX;
                    const int x = tx * 16 + ((pix - 1) & 15), y = ty * 16 + ((pix - 1) >> 4);
                    if (p.frameBuf) // the sample's own (frame, pass) image: added to the accumulator in order by k_sum_frames
                        p.frameBuf[(size_t)(frame * p.passes + (smp - 1) % p.passes) * ((size_t)p.W * p.H) + (x + (size_t)y * p.W)] = make_float4(L.x, L.y, L.z, 0);
                    else
                    {
                        float* a = (float*)(p.accum + (x + (size_t)y * p.W)); // renderer.cpp:124: accumulator += float4(sample, 0)
                        atomicAdd(a + 0, L.x), atomicAdd(a + 1, L.y), atomicAdd(a + 2, L.z);
                    }
                }
                // the slot's last sample: the tile's 256 x passes-th, or (one slot per pixel) the pixel's passes-th
                if (smp < (p.seedMode == RT_SEED_PER_PIXEL ? (px0 + 1) * p.passes : 256 * p.passes))
                {
                    float3 gD;
                    pt_generate(p, cam, slot, smp / p.passes, seed, gD);
                    p.rayO[slot] = make_float4(cam.pos.x, cam.pos.y, cam.pos.z, 0);
                    p.rayD[slot] = make_float4(gD.x, gD.y, gD.z, __int_as_float(0));
                    p.pix[slot] = smp + 1;
                }
                else alive = false;
            }
            p.seed[slot] = seed;
        }
        // compaction: one atomic per warp (ballot + popc), lanes write their rank
        const unsigned mask = __ballot_sync(0xffffffffu, alive);
        if (mask)
        {
            const int leader = __ffs(mask) - 1;
            int pos = 0;
            if (lane == leader) pos = atomicAdd(&p.count[cur ^ 1], __popc(mask));
            pos = __shfl_sync(0xffffffffu, pos, leader);
            if (alive) nextActive[pos + __popc(mask & ((1u << lane) - 1))] = slot;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Stream schedule: one persistent kernel, no global lock-step.
//
// With the reference's RNG a (tile, frame) stream is a serial chain of up to 256 x (depthLimit + 1)
// rays.  The wavefront advances every stream by one ray per extend/shade launch pair, so a render
// costs (longest chain) x (slowest ray of the iteration + shade + two launch gaps, cold L1 each
// launch): profiles/r1_v2_pt_per_iteration_persistent.txt shows 1536 iterations that never drop below
// ~70 us although the last 1000 of them carry a quarter of the rays.  Here every lane owns one stream
// and runs it to completion in registers (ray, seed, pixel counter, throughput factors); a lane whose
// stream ends pulls the next one (ballot / popc, one atomic per warp).  Streams are handed out tile by
// tile, so the 32 lanes of a warp trace the same tile in 32 different frames: similar rays, warm L1.
// ---------------------------------------------------------------------------------------------
constexpr int STREAM_MAX_DEPTH = 8;

// Pilot for the stream schedule: a stream's length is only known when it ends, and a long stream that
// is handed out late becomes the tail of the whole render (a chain of ~1500 dependent rays).  Sixteen
// throw-away paths per tile (own seeds, nothing is accumulated) estimate each tile's cost in node
// visits + triangle tests + a per-ray constant; the host sorts tiles by it and the stream pool hands
// them out longest first (LPT list scheduling).
constexpr int PILOT_PATHS = 16;

template <int ACCEL>
__global__ void __launch_bounds__(128) k_pt_pilot(const PTState p, const DScene s, const DCamera cam, unsigned int* __restrict__ cost)
{
    const int n = p.nTiles * PILOT_PATHS;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int k = i / PILOT_PATHS, j = i - k * PILOT_PATHS;
        const int tile = p.tileBegin + k * p.tileStep;
        const int tx = tile % p.tilesX, ty = tile / p.tilesX;
        uint32_t seed = wang_hash((uint32_t)i * 2654435761u + 12345u) | 1u;
        float3 O = cam.pos;
        float3 D = primary_dir(cam, (float)(tx * 16 + (j & 3) * 4) + 4 * random_float(seed), (float)(ty * 16 + (j >> 2) * 4) + 4 * random_float(seed));
        bool inside = false;
        unsigned int c = 0;
        for (int depth = 0;; depth++)
        {
            HitRec h;
            find_nearest<true, ACCEL>(s, O, D, 1e34f, h);
            c += 8u + (unsigned)h.traversed + 2u * (unsigned)h.tested;
            float3 L, w, nO, nD;
            bool nInside;
            if (pt_bounce(s, p.eps, p.depthLimit, O, D, inside, depth, h.t, h.u, h.v, h.obj, h.tri, seed, L, w, nO, nD, nInside)) break;
            O = nO, D = nD, inside = nInside;
        }
        atomicAdd(&cost[k], c);
    }
}

// Stream kernels: per-lane state machine with warp-level action voting.
//
// ncu on version 1 (profiles/r1_v3_k_pt_streams_v1_ncu_full.txt): issue slots 69 % busy but only 7.9 of
// 32 lanes active per instruction, because a warp waits for its slowest ray before it shades and lanes
// sit out each other's node / triangle / shading code.  Here every lane is in one of the states
//   NODE  next step is an interior-node visit          LEAF  next step is a triangle leaf, instance entry or exit
//   SHADE traversal finished, the bounce is pending    DEAD  no stream (pool exhausted)
// and each loop iteration the warp executes the ONE action most lanes are waiting for (ballot / popc),
// so the lanes that take part in an instruction are a majority instead of the leftovers.  Per lane the
// order of node visits, triangle tests, RNG draws and bounces is unchanged.
// (Version 2, the first voted kernel, was removed in round 2: profiles/r1_stream_kernel_* keep its numbers.)
enum { ST_DEAD = 0, ST_NODE = 1, ST_LEAF = 2, ST_SHADE = 3 };

// Stream kernel for FileScene's other accelerators (KD-tree, uniform grid): the same schedule as version 2 - one
// (tile, frame) RNG stream per lane, path state in registers, lanes pull the next stream when theirs ends - with
// the traversal expressed through the accelerator's cursor (rt_device.cuh KdCursor / GridCursor).  Three live states:
// TRAV (one node / one cell per step), TRI (one triangle of a leaf's / cell's run per step) and SHADE; the warp runs the
// action most lanes wait for and keeps repeating a TRAV or TRI action without a new vote while at least 3/4 of the lanes
// that entered it are still in that state.
enum { SA_DEAD = 0, SA_TRAV = 1, SA_SHADE = 2, SA_TRI = 3 };

template <int ACCEL>
__global__ void __launch_bounds__(128) k_pt_streams_alt(const PTState p, const DScene s, const DCamera cam,
    const int* __restrict__ tileOrder, const int frames, int* __restrict__ streamCounter)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int total = p.slots;
    bool poolEmpty = false;
    int state = SA_DEAD;
    int tile = 0, pix = 0, depth = 0;
    bool inside = false;
    uint32_t seed = 0;
    float3 wO = f3(0, 0, 0), wD = f3(0, 0, 0);
    float3 wst[STREAM_MAX_DEPTH];
    typename CursorOf<ACCEL>::type cursor;
    HitRec hit;
    hit.t = 0, hit.u = 0, hit.v = 0, hit.obj = -1, hit.tri = -1, hit.traversed = 0, hit.tested = 0;
    unsigned long long rays = 0;
    int frameIdx = 0; // frame of the launch this lane's stream belongs to (its samples go to that frame's images)

    // FindNearest prologue for the ray (wO, wD): light quad, floor plane (file_scene.cpp:172-173), then the accelerator
#define RT_START_RAY()                                                                         \
    {                                                                                          \
        hit.t = 1e34f, hit.u = 0, hit.v = 0, hit.obj = -1, hit.tri = -1;                       \
        float tq;                                                                              \
        if (quad_test(s, wO, wD, hit.t, tq)) hit.t = tq, hit.obj = 0;                          \
        const float3 fn = f3(s.floor_n[0], s.floor_n[1], s.floor_n[2]);                        \
        const float tp = -(dot(wO, fn) + s.floor_d) / (dot(wD, fn));                           \
        if (tp < hit.t && tp > 0) hit.t = tp, hit.obj = 1;                                     \
        state = cursor.start(s, wO, wD, hit, 0, -1) ? SA_SHADE : SA_TRAV;                      \
        rays++;                                                                                \
    }

    while (true)
    {
        const unsigned mDead = __ballot_sync(FULL, state == SA_DEAD);
        if (mDead && !poolEmpty)
        {
            const int nIdle = __popc(mDead);
            const int leader = __ffs(mDead) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(streamCounter, nIdle);
            base = __shfl_sync(FULL, base, leader);
            if (base + nIdle >= total) poolEmpty = true;
            const int stream = base + __popc(mDead & ((1u << lane) - 1));
            if (state == SA_DEAD && stream < total)
            {
                const int k = stream / frames, frame = stream - k * frames;
                tile = p.tileBegin + (tileOrder ? tileOrder[k] : k) * p.tileStep;
                seed = pt_seed(p, tile, p.firstSpp + frame * p.stride);
                pix = 0, depth = 0, inside = false, frameIdx = frame;
                const int tx = tile % p.tilesX, ty = tile / p.tilesX;
                const float jy = random_float(seed), jx = random_float(seed);
                wD = primary_dir(cam, (float)(tx * 16) + jx, (float)(ty * 16) + jy);
                wO = cam.pos;
                RT_START_RAY();
            }
        }
        const unsigned mTrav = __ballot_sync(FULL, state == SA_TRAV);
        const unsigned mTri = __ballot_sync(FULL, state == SA_TRI);
        const unsigned mShade = __ballot_sync(FULL, state == SA_SHADE);
        if ((mTrav | mTri | mShade) == 0) break;
        const int nT = __popc(mTrav), nR = __popc(mTri), nS = __popc(mShade);
        if (nT >= nR && nT >= nS)
        {
            const int keep = (nT * 3 + 3) >> 2;
            do
            {
                if (state == SA_TRAV)
                {
                    const int r = cursor.template step<false>(s, wO, wD, hit);
                    state = r == CUR_DONE ? SA_SHADE : (r == CUR_RUN ? SA_TRI : SA_TRAV);
                }
            } while (__popc(__ballot_sync(FULL, state == SA_TRAV)) >= keep);
        }
        else if (nR >= nS)
        {
            const int keep = (nR * 3 + 3) >> 2;
            do
            {
                if (state == SA_TRI)
                {
                    const int r = cursor.template tri_step<false, false>(s, wO, wD, hit);
                    state = r == CUR_DONE ? SA_SHADE : (r == CUR_RUN ? SA_TRI : SA_TRAV);
                }
            } while (__popc(__ballot_sync(FULL, state == SA_TRI)) >= keep);
        }
        else if (state == SA_SHADE)
        {
            float3 L, w, nO, nD;
            bool nInside;
            if (!pt_bounce(s, p.eps, p.depthLimit, wO, wD, inside, depth, hit.t, hit.u, hit.v, hit.obj, hit.tri, seed, L, w, nO, nD, nInside))
            {
                wst[depth] = w;
                depth++, wO = nO, wD = nD, inside = nInside;
                RT_START_RAY();
            }
            else
            {
                for (int d = depth - 1; d >= 0; d--) L = wst[d] * L;
                const int tx = tile % p.tilesX, ty = tile / p.tilesX;
                const int px = pix / p.passes; // pix counts samples: `passes` consecutive ones per pixel
                const int x = tx * 16 + (px & 15), y = ty * 16 + (px >> 4);
                if (p.frameBuf) // the sample's own (frame, pass) image: added to the accumulator in order by k_sum_frames
                    p.frameBuf[(size_t)(frameIdx * p.passes + pix % p.passes) * ((size_t)p.W * p.H) + (x + (size_t)y * p.W)] = make_float4(L.x, L.y, L.z, 0);
                else
                {
                    float* a = (float*)(p.accum + (x + (size_t)y * p.W));
                    atomicAdd(a + 0, L.x), atomicAdd(a + 1, L.y), atomicAdd(a + 2, L.z);
                }
                pix++;
                if (pix < 256 * p.passes)
                {
                    const int npx = pix / p.passes;
                    const int nx = tx * 16 + (npx & 15), ny = ty * 16 + (npx >> 4);
                    const float jy = random_float(seed), jx = random_float(seed);
                    wD = primary_dir(cam, (float)nx + jx, (float)ny + jy);
                    wO = cam.pos, depth = 0, inside = false;
                    RT_START_RAY();
                }
                else state = SA_DEAD;
            }
        }
    }
#undef RT_START_RAY
    for (int off = 16; off; off >>= 1) rays += __shfl_xor_sync(FULL, rays, off);
    if (lane == 0) atomicAdd(p.counters, rays);
}


// Stream kernel, version 5 = version 2 (one stream per lane, state in registers, ballot vote) with what the
// source-level profile of version 2 asked for (tools/ncu_source_hot.py on profiles/r1_v4_*):
//   * MISS is voted separately from surface shading: sky lookups (atan2f / acosf / texel) no longer
//     serialise against hit_info + lobe code inside one SHADE action (shading ran at 6.9 of 32 lanes);
//   * one copy of "finish the sample, next pixel" and one copy of the FindNearest prologue, behind flags,
//     instead of one per branch; the tile origin is kept in a register instead of tile % tilesX per sample;
//   * a NODE action keeps stepping without a re-vote while >= 3/4 of the lanes that entered it are still in
//     NODE (the vote was 10 % of all warp instructions), and the step itself selects instead of branching
//     (the stack pop ran at 2.2 lanes behind its own BSSY / BSYNC pair);
//   * every finished stream adds its duration (clock64) to its tile's cost: the next render call of the same
//     view sorts tiles by MEASURED cost instead of the 16-path pilot estimate (longest-first hand-out).
// Three other designs were built, parity-checked and measured slower, then removed (numbers in
// profiles/r1_stream_kernel_*.txt): six REDUX-voted states (more vote rounds per ray), K streams per lane in
// local memory (lanes 13.7 -> 17.4 but the parked state thrashes L1: hit rate 83 -> 64 %), and a per-warp pool
// of 64 streams in shared memory with traverse-only lanes and batched shading (lanes 17.3, but only 20 warps
// per SM fit and every warp's iteration got longer).  The kernel is bound by per-warp instruction latency x
// SIMD efficiency; throughput grows almost linearly with resident warps.
enum { ST_MISS = 4 };

template <bool TLAS, int MINB, bool PERPIXEL>
__global__ void __launch_bounds__(128, MINB) k_pt_streams5(const PTState p, const DScene s, const DCamera cam,
    const int* __restrict__ tileOrder, const int frames, int* __restrict__ streamCounter, unsigned long long* __restrict__ tileCost, const int keepShift, const unsigned laneMask)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int total = p.slots;
    const float4* __restrict__ nodes = s.nodes;
    const float4* __restrict__ tris = s.tris;
    bool poolEmpty = false;
    int state = ST_DEAD;
    // stream
    int tileXY = 0, pix = 0, depth = 0;
    bool inside = false;
    uint32_t seed = 0;
    unsigned int t0 = 0; // clock() at stream start (32 bit: a stream lasts milliseconds)
    float3 wO = f3(0, 0, 0), wD = f3(0, 0, 0); // the ray in world space
    float3 wst[STREAM_MAX_DEPTH];
    // traversal
    float3 O = f3(0, 0, 0), D = f3(0, 0, 0); // the ray in the space being traversed (TLAS only; flat: wO, wD)
    RaySlab rs = make_ray_slab(wO, wO);
    bool exact = false;
    int stack[STACK_SIZE];
    int sp = 0, cur = 0, instObj = -1;
    float ht = 0, hu = 0, hv = 0;
    int hobj = -1, htri = -1;
    unsigned int rays = 0;
#define TO (TLAS ? O : wO)
#define TD (TLAS ? D : wD)

    while (true)
    {
        const unsigned mNode = __ballot_sync(FULL, state == ST_NODE);
        const unsigned mLeaf = __ballot_sync(FULL, state == ST_LEAF);
        const unsigned mShade = __ballot_sync(FULL, state == ST_SHADE);
        const unsigned mMiss = __ballot_sync(FULL, state == ST_MISS);
        const unsigned mLive = mNode | mLeaf | mShade | mMiss;
        bool start = false;
        // (per-pixel streams end after every path: refill in batches of >= 8 lanes so that refills do not alternate with actions)
        if ((~mLive & laneMask) != 0 && !poolEmpty && (!PERPIXEL || mLive == 0 || __popc(~mLive & laneMask) >= 8))
        {
            // refill the dead lanes from the stream pool: one atomic per warp.  laneMask caps the streams per warp
            // for small jobs (fewer streams than lanes): a chain runs faster the fewer neighbours it waits for
            const unsigned mDead = ~mLive & laneMask;
            const int nIdle = __popc(mDead);
            const int leader = __ffs(mDead) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(streamCounter, nIdle);
            base = __shfl_sync(FULL, base, leader);
            if (base + nIdle >= total) poolEmpty = true;
            const int stream = base + __popc(mDead & ((1u << lane) - 1));
            if (state == ST_DEAD && ((laneMask >> lane) & 1) && stream < total)
            {
                // RT_SEED_REFERENCE_TILE: a stream is a (tile, frame) pair and runs the tile's 256 pixels; RT_SEED_PER_PIXEL:
                // a stream is ONE pixel of a (tile, frame) pair (32 consecutive streams = two pixel rows of one tile)
                const bool perPixel = PERPIXEL;
                const int unit = perPixel ? stream >> 8 : stream, px0 = perPixel ? stream & 255 : 0;
                const int k = unit / frames, frame = unit - k * frames;
                const int tile = p.tileBegin + (tileOrder ? tileOrder[k] : k) * p.tileStep;
                const int tx = tile % p.tilesX, ty = tile / p.tilesX;
                const int x = tx * 16 + (px0 & 15), y = ty * 16 + (px0 >> 4);
                seed = perPixel ? pt_pixel_seed(p, x, y, p.firstSpp + frame * p.stride) : pt_seed(p, tile, p.firstSpp + frame * p.stride);
                tileXY = (tx * 16) | ((ty * 16) << 16);
                pix = (px0 * p.passes) | (frame << 12), depth = 0, inside = false; // pix = sample of the tile (12 bits: 256 x passes <= 2048) | frame of the launch << 12
                const float jy = random_float(seed), jx = random_float(seed);
                wD = primary_dir(cam, (float)x + jx, (float)y + jy);
                wO = cam.pos;
                t0 = (unsigned int)clock();
                start = true;
            }
        }
        else
        {
            if (mLive == 0) break;
            const int nN = __popc(mNode), nL = __popc(mLeaf), nS = __popc(mShade), nM = __popc(mMiss);
            if (nN >= nL && nN >= nS && nN >= nM)
            {
                const int keep = nN - (nN >> keepShift);
                // the NaN-exact slab variant is chosen per ACTION, not per lane and box: if any lane of this action
                // holds a degenerate ray everyone takes the select-based min / max (same values for ordinary rays)
                bool inNode = state == ST_NODE, end = false;
                if (__any_sync(FULL, inNode && exact))
                {
                    do
                    {
                        if (inNode) node_step<true>(nodes, rs, ht, stack, sp, cur, end), inNode = !end && cur >= 0;
                    } while (__popc(__ballot_sync(FULL, inNode)) >= keep);
                }
                else
                {
                    do
                    {
                        if (inNode) node_step<false>(nodes, rs, ht, stack, sp, cur, end), inNode = !end && cur >= 0;
                    } while (__popc(__ballot_sync(FULL, inNode)) >= keep);
                }
                if (state == ST_NODE) state = end ? (hobj == -1 ? ST_MISS : ST_SHADE) : (cur >= 0 ? ST_NODE : ST_LEAF);
            }
            else if (nL >= nS && nL >= nM)
            {
                if (state == ST_LEAF)
                {
                    const int payload = ~cur;
                    bool pop = true;
                    if (TLAS && payload == SENTINEL_PAYLOAD)
                    {
                        O = wO, D = wD, rs = make_ray_slab(wO, recip(wD)), exact = needs_exact_slab(wO, wD); // blas_bvh.cpp:385-388
                    }
                    else if (TLAS && (payload & INSTANCE_BIT))
                    {
                        const float4* I = s.inst + 4 * (size_t)(payload & ~INSTANCE_BIT);
                        const float4 r0 = __ldg(I), r1 = __ldg(I + 1), r2 = __ldg(I + 2);
                        const int4 meta = __ldg((const int4*)(I + 3));
                        O = f3((wO.x * r0.x + wO.y * r0.y) + (wO.z * r0.z + r0.w),
                               (wO.x * r1.x + wO.y * r1.y) + (wO.z * r1.z + r1.w),
                               (wO.x * r2.x + wO.y * r2.y) + (wO.z * r2.z + r2.w));
                        D = f3((wD.x * r0.x + wD.y * r0.y) + wD.z * r0.z,
                               (wD.x * r1.x + wD.y * r1.y) + wD.z * r1.z,
                               (wD.x * r2.x + wD.y * r2.y) + wD.z * r2.z);
                        rs = make_ray_slab(O, recip(D)), exact = needs_exact_slab(O, D);
                        instObj = meta.y;
                        stack[sp++] = ~SENTINEL_PAYLOAD;
                        cur = meta.x, state = cur >= 0 ? ST_NODE : ST_LEAF;
                        pop = false;
                    }
                    else
                    {
                        int slot = payload;
                        while (true)
                        {
                            const float4* T = tris + 3 * (size_t)slot;
                            const float4 t0_ = __ldg(T), t1 = __ldg(T + 1), t2 = __ldg(T + 2);
                            const int tag = __float_as_int(t0_.w);
                            if (intersect_tri(TO, TD, f3(t0_.x, t0_.y, t0_.z), f3(t1.x, t1.y, t1.z), f3(t2.x, t2.y, t2.z), ht, hu, hv))
                            {
                                htri = tag & ~LAST_BIT;
                                hobj = instObj >= 0 ? instObj : __float_as_int(t1.w);
                            }
                            if (tag & LAST_BIT) break;
                            slot++;
                        }
                    }
                    if (pop)
                    {
                        if (sp == 0) state = hobj == -1 ? ST_MISS : ST_SHADE;
                        else cur = stack[--sp], state = cur >= 0 ? ST_NODE : ST_LEAF;
                    }
                }
            }
            else
            {
                // MISS (sky, renderer.cpp:54) or surface shading, whichever more lanes wait for; then ONE copy
                // of "sample finished -> splat, next pixel" for the lanes whose path ended
                const bool doMiss = nM >= nS;
                bool fin = false;
                float3 L = f3(0, 0, 0);
                if (doMiss)
                {
                    if (state == ST_MISS) L = sky_color(s, wD), fin = true;
                }
                else if (state == ST_SHADE)
                {
                    float3 w, I, N, nD;
                    bool nInside;
                    const int k = pt_surface(s, p.depthLimit, wO, wD, inside, depth, ht, hu, hv, hobj, htri, seed, L, w, I, N, nD, nInside);
                    if (k == PT_END) fin = true;
                    else
                    {
                        if (k == PT_DIFF)
                        {
                            nD = diffuse_reflection(N, seed);
                            w = w * dot(nD, N);
                        }
                        wst[depth] = w;
                        depth++, wO = I + nD * p.eps, wD = nD, inside = nInside;
                        start = true;
                    }
                }
                if (fin)
                {
                    for (int d = depth - 1; d >= 0; d--) L = wst[d] * L;
                    const int x0 = tileXY & 0xffff, y0 = tileXY >> 16;
                    const int px = (pix & 4095) / p.passes; // `passes` consecutive samples per pixel (renderer.cpp:123)
                    const size_t pixel = (x0 + (px & 15)) + (size_t)(y0 + (px >> 4)) * p.W;
                    if (p.frameBuf) p.frameBuf[(size_t)(pix >> 12) * ((size_t)p.W * p.H) + pixel] = make_float4(L.x, L.y, L.z, 0); // the frame's own image
                    else
                    {
                        float* a = (float*)(p.accum + pixel); // renderer.cpp:124
                        atomicAdd(a + 0, L.x), atomicAdd(a + 1, L.y), atomicAdd(a + 2, L.z);
                    }
                    pix++;
                    if (PERPIXEL ? (pix & 4095) % p.passes != 0 : (pix & 4095) < 256 * p.passes)
                    {
                        const int nx = (pix & 4095) / p.passes;
                        const float jy = random_float(seed), jx = random_float(seed);
                        wD = primary_dir(cam, (float)(x0 + (nx & 15)) + jx, (float)(y0 + (nx >> 4)) + jy);
                        wO = cam.pos, depth = 0, inside = false;
                        start = true;
                    }
                    else
                    {
                        state = ST_DEAD;
                        if (tileCost) atomicAdd(&tileCost[((y0 >> 4) * p.tilesX + (x0 >> 4) - p.tileBegin) / p.tileStep], (unsigned long long)((unsigned int)clock() - t0));
                    }
                }
            }
        }
        if (start)
        {
            // FindNearest prologue for the ray (wO, wD): light quad, floor plane (file_scene.cpp:172-173), then the BVH
            ht = 1e34f, hu = 0, hv = 0, hobj = -1, htri = -1;
            float tq;
            if (quad_test(s, wO, wD, ht, tq)) ht = tq, hobj = 0;
            const float3 fn = f3(s.floor_n[0], s.floor_n[1], s.floor_n[2]);
            const float tp = -(dot(wO, fn) + s.floor_d) / (dot(wD, fn));
            if (tp < ht && tp > 0) ht = tp, hobj = 1;
            if (TLAS) O = wO, D = wD;
            rs = make_ray_slab(wO, recip(wD)), exact = needs_exact_slab(wO, wD);
            sp = 0, cur = s.root_ref, instObj = s.flat_obj_idx;
            state = cur >= 0 ? ST_NODE : ST_LEAF;
            rays++;
        }
    }
#undef TO
#undef TD
    for (int off = 16; off; off >>= 1) rays += __shfl_xor_sync(FULL, rays, off);
    if (lane == 0) atomicAdd(p.counters, (unsigned long long)rays);
}

} // namespace rtb

#include "rt_streams8.cuh"

namespace rtb {

// ---------------------------------------------------------------------------------------------
// Whitted wavefront
// ---------------------------------------------------------------------------------------------
struct WhState {
    float4* rayO[2];   // O.xyz, -
    float4* rayD[2];   // D.xyz, int flags = inside | depth << 8
    float4* rayW[2];   // weight.xyz, int pixel
    float4* hit;       // t, u, v, int objIdx
    int* hitTri;
    float4* shadow;    // 5 x float4 per entry: (O, tmax) (D, int pixel) (w~, -) (coef, -) (irradiance, -)
    int* count;        // [0],[1] ray queues, [2] shadow queue, [3] overflow flag
    unsigned long long* counters;
    float4* accum;
    int capacity, W, H, depthLimit;
    float eps;
};

__global__ void __launch_bounds__(256) k_wh_generate(const WhState p, const DCamera* __restrict__ camPtr)
{
    const DCamera cam = *camPtr; // in device memory so that the captured frame graph can be replayed with a new camera
    const int n = p.W * p.H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int x = i % p.W, y = i / p.W;
        const float3 D = primary_dir(cam, (float)x, (float)y); // renderer.cpp:144
        p.rayO[0][i] = make_float4(cam.pos.x, cam.pos.y, cam.pos.z, 0);
        p.rayD[0][i] = make_float4(D.x, D.y, D.z, __int_as_float(0));
        p.rayW[0][i] = make_float4(1, 1, 1, __int_as_float(i));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) p.count[0] = n, p.count[1] = 0, p.count[2] = 0, p.count[4] = 0, p.count[5] = 0;
}

template <int ACCEL>
__global__ void __launch_bounds__(128) k_wh_extend(const WhState p, const DScene s, int cur)
{
    const int n = min(p.count[cur], p.capacity);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float4 o = p.rayO[cur][i], d = p.rayD[cur][i];
        HitRec h;
        find_nearest<false, ACCEL>(s, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), 1e34f, h);
        p.hit[i] = make_float4(h.t, h.u, h.v, __int_as_float(h.obj));
        p.hitTri[i] = h.tri;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        p.count[cur ^ 1] = 0, p.count[2] = 0;
        p.counters[0] += (unsigned long long)n;
        p.counters[2] += 1;
    }
}

struct WhSrc {
    const WhState& p;
    int cur;
    __device__ __forceinline__ bool load(int i, float3& O, float3& D, float& tmax) const
    {
        const float4 o = p.rayO[cur][i], d = p.rayD[cur][i];
        O = f3(o.x, o.y, o.z), D = f3(d.x, d.y, d.z), tmax = 1e34f;
        return true;
    }
    __device__ __forceinline__ void world(int i, float3& O, float3& D) const
    {
        const float4 o = p.rayO[cur][i], d = p.rayD[cur][i];
        O = f3(o.x, o.y, o.z), D = f3(d.x, d.y, d.z);
    }
    __device__ __forceinline__ void store(int i, const HitRec& h) const
    {
        p.hit[i] = make_float4(h.t, h.u, h.v, __int_as_float(h.obj));
        p.hitTri[i] = h.tri;
    }
};

template <int ACCEL>
__global__ void __launch_bounds__(128) k_wh_extend_persistent(const WhState p, const DScene s, int cur)
{
    const int n = min(p.count[cur], p.capacity);
    WhSrc src = { p, cur };
    accel_trace_queue<ACCEL, false, false>(s, src, n, p.count + 4);
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        p.count[cur ^ 1] = 0, p.count[2] = 0, p.count[5] = 0;
        p.counters[0] += (unsigned long long)n;
        p.counters[2] += 1;
    }
}

__device__ __forceinline__ void splat(float4* accum, int pixel, float3 c)
{
    float* a = (float*)(accum + pixel);
    atomicAdd(a + 0, c.x), atomicAdd(a + 1, c.y), atomicAdd(a + 2, c.z);
}

// warp-aggregated queue append; returns the slot or -1 (not pushing / overflow)
__device__ __forceinline__ int queue_push(bool want, int* counter, int capacity, int* overflow)
{
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (!mask) return -1;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int pos = 0;
    if (lane == leader) pos = atomicAdd(counter, __popc(mask));
    pos = __shfl_sync(0xffffffffu, pos, leader);
    if (!want) return -1;
    pos += __popc(mask & ((1u << lane) - 1));
    if (pos >= capacity) { *overflow = 1; return -1; }
    return pos;
}

// One level of Renderer::Trace (2. WhittedStyle/renderer.cpp:21-91), top-down: a ray's weight is the
// product of the factors the reference applies on the way back up (medium_scale, reflectivity*albedo,
// albedo*(1-Fr), albedo*Fr); the diffuse term keeps the reference's own grouping
// diffuseness*brdf*(irradiance + ambient) by deferring it to the connect stage.
__global__ void __launch_bounds__(128) k_wh_shade(const WhState p, const DScene s, int cur)
{
    const int n = min(p.count[cur], p.capacity);
    const int nxt = cur ^ 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.count[4] = 0; // fetch counter of the next extend launch
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x)
    {
        const int i = base + threadIdx.x;
        bool push1 = false, push2 = false, pushS = false;
        float3 o1, d1, w1, o2, d2, w2, so, sd, sw, scoef, sirr;
        float stmax = 0;
        int f1 = 0, f2 = 0, pixel = 0;
        if (i < n)
        {
            const float4 o4 = p.rayO[cur][i], d4 = p.rayD[cur][i], w4 = p.rayW[cur][i], h4 = p.hit[i];
            const float3 O = f3(o4.x, o4.y, o4.z), D = f3(d4.x, d4.y, d4.z), w = f3(w4.x, w4.y, w4.z);
            pixel = __float_as_int(w4.w);
            const int flags = __float_as_int(d4.w);
            const bool inside = flags & 1;
            const int depth = flags >> 8;
            const int obj = __float_as_int(h4.w);
            const float t = h4.x;
            if (obj == -1) splat(p.accum, pixel, w * sky_color(s, D)); // renderer.cpp:25
            else
            {
                const float3 I = O + t * D;
                ShadeHit h;
                float uu, vv;
                hit_info(s, D, I, obj, p.hitTri[i], h4.y, h4.z, h, uu, vv);
                if (h.isLight) splat(p.accum, pixel, w * f3(s.light_color[0], s.light_color[1], s.light_color[2]));
                else
                {
                    float3 medium_scale = f3(1, 1, 1);
                    if (inside) // renderer.cpp:81-88
                        medium_scale = beer_scale(h.absorption.x * -t, h.absorption.y * -t, h.absorption.z * -t);
                    const float3 wm = w * medium_scale;
                    const float reflectivity = h.reflectivity, refractivity = h.refractivity;
                    const float diffuseness = 1 - (reflectivity + refractivity);
                    const bool deeper = depth + 1 <= p.depthLimit; // Trace returns 0 when depth > depthLimit (:23)
                    if (reflectivity > 0.0f)
                    {
                        const float3 R = reflect(D, h.N);
                        if (deeper) push1 = true, o1 = I + R * p.eps, d1 = R, w1 = wm * (reflectivity * h.albedo), f1 = (depth + 1) << 8;
                    }
                    else if (refractivity > 0.0f)
                    {
                        const float3 R = reflect(D, h.N);
                        const float n1 = inside ? 1.2f : 1, n2 = inside ? 1 : 1.2f;
                        const float eta = n1 / n2, cosi = dot(-D, h.N);
                        const float cost2 = 1.0f - eta * eta * (1 - cosi * cosi);
                        float Fr = 1;
                        if (cost2 > 0)
                        {
                            const float a = n1 - n2, b = n1 + n2, R0 = (a * a) / (b * b), c = 1 - cosi;
                            Fr = R0 + (1 - R0) * (c * c * c * c * c);
                            const float3 T = eta * D + ((eta * cosi - sqrtf(fabsf(cost2))) * h.N);
                            if (deeper) push2 = true, o2 = I + T * p.eps, d2 = T, w2 = wm * (h.albedo * (1 - Fr)), f2 = ((depth + 1) << 8) | (inside ? 0 : 1);
                        }
                        if (deeper) push1 = true, o1 = I + R * p.eps, d1 = R, w1 = wm * (h.albedo * Fr), f1 = (depth + 1) << 8;
                    }
                    if (diffuseness > 0)
                    {
                        // DirectIllumination renderer.cpp:105-126
                        const float3 ambient = f3(0.3f, 0.3f, 0.3f);
                        const float3 brdf = h.albedo * RT_INVPI;
                        const float3 coef = diffuseness * brdf;
                        float3 Lv = f3(s.light_pos[0], s.light_pos[1], s.light_pos[2]) - I;
                        const float distance = sqrtf(dot(Lv, Lv));
                        Lv = Lv * (1 / distance);
                        const float ndotl = dot(h.N, Lv);
                        if (ndotl < p.eps) splat(p.accum, pixel, wm * (coef * (f3(0, 0, 0) + ambient)));
                        else
                        {
                            const float attenuation = 1 / (distance * distance);
                            const float3 in_radiance = f3(s.light_color[0], s.light_color[1], s.light_color[2]) * attenuation;
                            pushS = true, so = I + Lv * p.eps, sd = Lv, stmax = distance - 2 * p.eps;
                            sw = wm, scoef = coef, sirr = in_radiance * dot(h.N, Lv);
                        }
                    }
                }
            }
        }
        // refraction ray first, reflection second (evaluation order of renderer.cpp:68-71; only the
        // accumulation order depends on it)
        int q = queue_push(push2, &p.count[nxt], p.capacity, &p.count[3]);
        if (q >= 0)
        {
            p.rayO[nxt][q] = make_float4(o2.x, o2.y, o2.z, 0);
            p.rayD[nxt][q] = make_float4(d2.x, d2.y, d2.z, __int_as_float(f2));
            p.rayW[nxt][q] = make_float4(w2.x, w2.y, w2.z, __int_as_float(pixel));
        }
        q = queue_push(push1, &p.count[nxt], p.capacity, &p.count[3]);
        if (q >= 0)
        {
            p.rayO[nxt][q] = make_float4(o1.x, o1.y, o1.z, 0);
            p.rayD[nxt][q] = make_float4(d1.x, d1.y, d1.z, __int_as_float(f1));
            p.rayW[nxt][q] = make_float4(w1.x, w1.y, w1.z, __int_as_float(pixel));
        }
        q = queue_push(pushS, &p.count[2], p.capacity, &p.count[3]);
        if (q >= 0)
        {
            float4* e = p.shadow + 5 * (size_t)q;
            e[0] = make_float4(so.x, so.y, so.z, stmax);
            e[1] = make_float4(sd.x, sd.y, sd.z, __int_as_float(pixel));
            e[2] = make_float4(sw.x, sw.y, sw.z, 0);
            e[3] = make_float4(scoef.x, scoef.y, scoef.z, 0);
            e[4] = make_float4(sirr.x, sirr.y, sirr.z, 0);
        }
    }
}

// connect: shadow rays in their own any-hit kernel (IsOccluded, file_scene.cpp:177-187), then
// out_radiance += diffuseness * brdf * (irradiance + ambient) (renderer.cpp:74-79)
template <int ACCEL>
__global__ void __launch_bounds__(128) k_wh_connect(const WhState p, const DScene s)
{
    const int n = min(p.count[2], p.capacity);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float4* e = p.shadow + 5 * (size_t)i;
        const float4 a = e[0], b = e[1], w = e[2], c = e[3], ir = e[4];
        const bool occluded = is_occluded<ACCEL>(s, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w);
        const float3 irradiance = occluded ? f3(0, 0, 0) : f3(ir.x, ir.y, ir.z);
        const float3 ambient = f3(0.3f, 0.3f, 0.3f);
        splat(p.accum, __float_as_int(b.w), f3(w.x, w.y, w.z) * (f3(c.x, c.y, c.z) * (irradiance + ambient)));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) p.counters[1] += (unsigned long long)n;
}

// connect, persistent-warp version: the shadow queue through trace_queue<ANYHIT>
struct ShadowSrc {
    const WhState& p;
    __device__ __forceinline__ bool load(int i, float3& O, float3& D, float& tmax) const
    {
        const float4* e = p.shadow + 5 * (size_t)i;
        const float4 a = e[0], b = e[1];
        O = f3(a.x, a.y, a.z), D = f3(b.x, b.y, b.z), tmax = a.w;
        return true;
    }
    __device__ __forceinline__ void world(int i, float3& O, float3& D) const
    {
        const float4* e = p.shadow + 5 * (size_t)i;
        const float4 a = e[0], b = e[1];
        O = f3(a.x, a.y, a.z), D = f3(b.x, b.y, b.z);
    }
    __device__ __forceinline__ void store(int i, const HitRec& h) const
    {
        const float4* e = p.shadow + 5 * (size_t)i;
        const float4 b = e[1], w = e[2], c = e[3], ir = e[4];
        const float3 irradiance = h.obj > -1 ? f3(0, 0, 0) : f3(ir.x, ir.y, ir.z);
        const float3 ambient = f3(0.3f, 0.3f, 0.3f);
        splat(p.accum, __float_as_int(b.w), f3(w.x, w.y, w.z) * (f3(c.x, c.y, c.z) * (irradiance + ambient)));
    }
};

template <int ACCEL>
__global__ void __launch_bounds__(128) k_wh_connect_persistent(const WhState p, const DScene s)
{
    const int n = min(p.count[2], p.capacity);
    ShadowSrc src = { p };
    accel_trace_queue<ACCEL, true, false>(s, src, n, p.count + 5);
    if (blockIdx.x == 0 && threadIdx.x == 0) p.counters[1] += (unsigned long long)n;
}

// screen->pixels: RGBF32_to_RGB8 (template/precomp.h:325-341, scalar branch) of accumulator * scale
__global__ void k_to_rgb8(const float4* __restrict__ accum, uint32_t* __restrict__ out, int n, float scale)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float4 a = accum[i];
        const float x = a.x * scale, y = a.y * scale, z = a.z * scale;
        const uint32_t r = (uint32_t)(255.0f * smin(1.0f, x));
        const uint32_t g = (uint32_t)(255.0f * smin(1.0f, y));
        const uint32_t b = (uint32_t)(255.0f * smin(1.0f, z));
        out[i] = (r << 16) + (g << 8) + b;
    }
}

} // namespace rtb

using namespace rtb;

static bool is_alt_kind(int kind) { return kind >= RT_SCENE_FLAT_KDTREE && kind <= RT_SCENE_TLAS_GRID; }

struct rt_renderer {
    rt_scene* scene = nullptr;
    rt_render_params params = {};
    DCamera cam = {};
    cudaStream_t ownStream = nullptr, stream = nullptr;
    float4* ownAccum = nullptr;
    float4* accum = nullptr;
    void* importedAccum = nullptr; // another process' accumulator opened through CUDA IPC (rt_renderer_import_accumulator)
    int sms = 148;
    // path tracer
    PTState pt = {};
    int ptSlotsAllocated = 0;
    int ptIterations = 0;
    bool persistent = true;
    bool useStreams = true;
    int passes = 1;   // Renderer::passes (3. PathTracer/renderer.h:50), changeable between frames like the UI slider does
    int streamCtasPerSm = 8;
    int streamKernel = 8; // 8 = current (rt_streams8.cuh); 5 = the round-1 kernel, kept for A/B profiling (RT_B200_STREAM_KERNEL)
    bool streamMeasuredLpt = true;
    int streamMinB = 7, streamKeepShift = 1;
    // version 8 without RT_B200_STREAM_MINB: 8 CTAs/SM (64 registers, shading code spills) for throughput-bound jobs - TLAS scenes, or
    // >= 8 streams per resident lane - and 7 (72 registers) where the longest chain bounds the job (profiles/r2_stream_kernel_sweeps.txt block 8)
    bool streamMinBAuto = false;
    int streamCtasPerSm8 = 8;
    bool streamFastNode = true; // version 8: skip the full vote while interior-node lanes are the majority
    int streamSmemSlots = 0;  // version 8: stack slots per thread in shared memory (0 = local-memory stack), from the scene's tree depths
    // ordered accumulation: every sample of a launch goes to its own (frame, pass) image, k_sum_frames adds them in order
    bool orderedFrames = true;
    float4* dImages = nullptr;
    size_t imagesCapacity = 0;            // in pixels (float4 each)
    size_t imageBudgetBytes = 8ull << 30; // frames per launch = budget / bytes per frame (RT_B200_IMAGE_BUDGET_MB)
    // look-ahead (rt_render_params.lookahead_frames): frames rendered ahead of the Tick sequence
    float4* dFrameBuf = nullptr;
    int aheadCapacity = 0, aheadBase = 0, aheadStride = 1, aheadReady = 0;
    bool aheadValid = false;
    bool streamLaneCap = true;
    bool streamLpt = true;
    int* dTileOrder = nullptr;
    unsigned int* dTileCost = nullptr;
    int tileOrderCapacity = 0, tileOrderCount = 0;
    bool tileOrderValid = false;
    unsigned long long* dTileClock = nullptr; // per-tile sum of stream durations of the last stream-kernel launch
    int tileOrderSource = 0;                  // 0 none, 1 pilot estimate, 2 measured durations
    bool tileClockRecorded = false;
    int lastStreamFrames = 1;
    std::vector<void*> allocations;
    // whitted
    WhState wh = {};
    DCamera* dCam = nullptr;   // camera in device memory (read by k_wh_generate)
    bool whGraphEnabled = true;
    cudaGraphExec_t whGraphExec = nullptr;
    float4* whGraphAccum = nullptr;
    int whGraphLaunches = 0;
    DCamera whLastCam = {};      // camera of the last Whitted frame (an overflowed frame is rendered again by rt_renderer_sync)
    bool whFramePending = false; // a Whitted frame has been launched and its overflow flag not yet checked
    // counters
    unsigned long long* dCounters = nullptr; // 4
    int* dCount = nullptr;                   // 4
    int* hCount = nullptr;                   // pinned
    uint64_t paths = 0, launches = 0;
    uint32_t* dPixels = nullptr;
    bool overflowed = false;
    // per-stage profiling (rt_renderer_set_profiling)
    bool profiling = false;
    std::vector<cudaEvent_t> evPool;
    size_t evUsed = 0;
    struct Span { int stage; cudaEvent_t a, b; };
    std::vector<Span> spans;
    cudaEvent_t openEvent = nullptr;

    void prof_begin()
    {
        if (!profiling) return;
        openEvent = next_event();
        cudaEventRecord(openEvent, stream);
    }
    void prof_end(int stage)
    {
        launches++;
        if (!profiling) return;
        cudaEvent_t b = next_event();
        cudaEventRecord(b, stream);
        spans.push_back({ stage, openEvent, b });
    }
    cudaEvent_t next_event()
    {
        if (evUsed == evPool.size())
        {
            cudaEvent_t e;
            cudaEventCreate(&e);
            evPool.push_back(e);
        }
        return evPool[evUsed++];
    }
};


// stream kernel version 5: dispatch on (TLAS, min CTAs per SM the register budget is bounded for)
typedef void (*Streams5Fn)(const PTState, const DScene, const DCamera, const int*, const int, int*, unsigned long long*, const int, const unsigned);
template <bool PERPIXEL>
static Streams5Fn streams5_kernel_for(bool tlas, int minb)
{
    if (tlas) return minb <= 6 ? k_pt_streams5<true, 6, PERPIXEL> : minb >= 8 ? k_pt_streams5<true, 8, PERPIXEL> : k_pt_streams5<true, 7, PERPIXEL>;
    return minb <= 6 ? k_pt_streams5<false, 6, PERPIXEL> : minb >= 8 ? k_pt_streams5<false, 8, PERPIXEL> : k_pt_streams5<false, 7, PERPIXEL>;
}
// the seed mode is a template argument so that the reference-RNG kernel carries none of the per-pixel stream code
static Streams5Fn streams5_kernel(bool tlas, int minb, bool perPixel = false)
{
    return perPixel ? streams5_kernel_for<true>(tlas, minb) : streams5_kernel_for<false>(tlas, minb);
}

// stream kernel version 8: dispatch on (TLAS, min CTAs per SM, seed mode, shared-memory stack slots)
template <bool TLAS, bool PERPIXEL>
static Streams5Fn streams8_kernel_for(int minb, int slots)
{
    if (PERPIXEL) return slots >= 32 ? k_pt_streams8<TLAS, 7, PERPIXEL, 32> : k_pt_streams8<TLAS, 7, PERPIXEL, 0>;
    if (minb >= 8) return slots >= 32 ? k_pt_streams8<TLAS, 8, PERPIXEL, 32> : slots >= 24 ? k_pt_streams8<TLAS, 8, PERPIXEL, 24> : k_pt_streams8<TLAS, 8, PERPIXEL, 0>;
    return slots >= 32 ? k_pt_streams8<TLAS, 7, PERPIXEL, 32> : slots >= 24 ? k_pt_streams8<TLAS, 7, PERPIXEL, 24> : k_pt_streams8<TLAS, 7, PERPIXEL, 0>;
}
static Streams5Fn streams8_kernel(bool tlas, int minb, bool perPixel, int slots)
{
    if (tlas) return perPixel ? streams8_kernel_for<true, true>(minb, slots) : streams8_kernel_for<true, false>(minb, slots);
    return perPixel ? streams8_kernel_for<false, true>(minb, slots) : streams8_kernel_for<false, false>(minb, slots);
}
// shared-memory stack slots for a scene: its deepest possible stack + the CUR_END slot, rounded up to a compiled size; 0 = local memory
static int streams8_slots(const rt_scene* sc, bool perPixel)
{
    const int need = sc->stack_entries + 1;
    if (const char* e = getenv("RT_B200_STREAM_SMEM_SLOTS")) { const int v = atoi(e); if (v == 0 || (v >= need && (v == 24 || v == 32))) return (perPixel && v == 24) ? 32 : v; }
    // default: the local-memory stack.  Same-box A/B (profiles/r2_stream_kernel_sweeps.txt): shared-memory columns are 1-4 % slower -
    // the carve-out (12-16 KB per CTA) comes out of L1, whose hit rate falls from 84 % to 75 % on the bench scene.
    (void)perPixel;
    return 0;
}

template <class T>
static rt_status ralloc(rt_renderer* r, T** p, size_t bytes)
{
    *p = nullptr;
    RT_CUDA(cudaMalloc((void**)p, bytes ? bytes : 16));
    r->allocations.push_back(*p);
    return RT_OK;
}

static void rfree(rt_renderer* r, void* p)
{
    if (!p) return;
    for (size_t i = 0; i < r->allocations.size(); i++)
        if (r->allocations[i] == p) { r->allocations.erase(r->allocations.begin() + i); break; }
    cudaFree(p);
}

// (re)allocates the Whitted ray / hit / shadow queues for `capacity` entries; the frame graph captured for the old buffers is dropped
static rt_status whitted_alloc_queues(rt_renderer* r, size_t capacity)
{
    WhState& w = r->wh;
    if (capacity > (size_t)0x7fffffff) capacity = 0x7fffffff;
    if (r->stream) cudaStreamSynchronize(r->stream);
    if (r->whGraphExec) { cudaGraphExecDestroy(r->whGraphExec); r->whGraphExec = nullptr; }
    rt_status st;
    for (int k = 0; k < 2; k++)
    {
        rfree(r, w.rayO[k]), rfree(r, w.rayD[k]), rfree(r, w.rayW[k]);
        if ((st = ralloc(r, &w.rayO[k], capacity * 16)) != RT_OK) return st;
        if ((st = ralloc(r, &w.rayD[k], capacity * 16)) != RT_OK) return st;
        if ((st = ralloc(r, &w.rayW[k], capacity * 16)) != RT_OK) return st;
    }
    rfree(r, w.hit), rfree(r, w.hitTri), rfree(r, w.shadow);
    if ((st = ralloc(r, &w.hit, capacity * 16)) != RT_OK) return st;
    if ((st = ralloc(r, &w.hitTri, capacity * 4)) != RT_OK) return st;
    if ((st = ralloc(r, &w.shadow, capacity * 80)) != RT_OK) return st;
    w.capacity = (int)capacity;
    return RT_OK;
}

static DCamera make_camera(const rt_camera& c, int W, int H)
{
    DCamera d;
    d.pos = make_float3(c.pos[0], c.pos[1], c.pos[2]);
    d.topLeft = make_float3(c.top_left[0], c.top_left[1], c.top_left[2]);
    d.topRight = make_float3(c.top_right[0], c.top_right[1], c.top_right[2]);
    d.bottomLeft = make_float3(c.bottom_left[0], c.bottom_left[1], c.bottom_left[2]);
    d.invW = 1.0f / W, d.invH = 1.0f / H; // camera.h:26-27
    return d;
}

static int num_tiles(const rt_render_params& p)
{
    const int all = (p.width / 16) * (p.height / 16); // renderer.cpp:151 (integer division)
    const int end = p.tile_end > 0 ? (p.tile_end < all ? p.tile_end : all) : all;
    const int step = p.tile_step > 0 ? p.tile_step : 1;
    const int n = (end - p.tile_begin + step - 1) / step; // tiles tile_begin, tile_begin + step, ... < end
    return n > 0 ? n : 0;
}

extern "C" {

void rt_render_params_default(rt_render_params* p, int integrator, int width, int height)
{
    memset(p, 0, sizeof *p);
    p->integrator = integrator, p->width = width, p->height = height;
    p->depth_limit = 5, p->epsilon = 0.001f, p->seed_mode = RT_SEED_REFERENCE_TILE;
}

rt_status rt_renderer_create(rt_scene* scene, const rt_render_params* params, rt_renderer** out)
{
    if (!scene || !params || !out) { set_error("rt_renderer_create: null argument"); return RT_ERR_INVALID; }
    *out = nullptr;
    if (params->width <= 0 || params->height <= 0 || params->depth_limit < 0 || params->depth_limit > 64)
    {
        set_error("rt_renderer_create: bad width / height / depth_limit");
        return RT_ERR_INVALID;
    }
    if (params->integrator != RT_INTEGRATOR_PATH && params->integrator != RT_INTEGRATOR_WHITTED)
    {
        set_error("rt_renderer_create: unknown integrator");
        return RT_ERR_INVALID;
    }
    RT_CUDA(cudaSetDevice(scene->device));
    if (scene->destroy_requested.load()) { set_error("rt_renderer_create: the scene has been destroyed"); return RT_ERR_INVALID; }
    rt_renderer* r = new rt_renderer();
    r->scene = scene, r->params = *params;
    scene->renderers.fetch_add(1);
    r->passes = params->passes > 0 ? params->passes : 1;
    if (r->passes > 8) { set_error("rt_renderer_create: passes > 8 (the reference's slider stops at 4)"); delete r; return RT_ERR_UNSUPPORTED; }
    cudaDeviceGetAttribute(&r->sms, cudaDevAttrMultiProcessorCount, scene->device);
    rt_camera cam;
    rt_camera_default(&cam, params->width, params->height);
    r->cam = make_camera(cam, params->width, params->height);
    rt_status st = RT_OK;
    auto fail = [&](rt_status e) { rt_renderer_destroy(r); return e; };
    if (cudaStreamCreateWithFlags(&r->ownStream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream create failed"); return fail(RT_ERR_CUDA); }
    r->stream = r->ownStream;
    scene->apply_l2_policy(r->stream); // fat nodes persisting in L2 against the streaming frame images (RT_B200_L2_PERSIST_MB)
    const size_t px = (size_t)params->width * params->height;
    if ((st = ralloc(r, &r->ownAccum, px * 16)) != RT_OK) return fail(st);
    r->accum = r->ownAccum;
    if (cudaMemset(r->ownAccum, 0, px * 16) != cudaSuccess) return fail(RT_ERR_CUDA);
    if ((st = ralloc(r, &r->dCounters, 4 * sizeof(unsigned long long))) != RT_OK) return fail(st);
    if ((st = ralloc(r, &r->dCount, 8 * sizeof(int))) != RT_OK) return fail(st);
    cudaMemset(r->dCounters, 0, 4 * sizeof(unsigned long long));
    cudaMemset(r->dCount, 0, 8 * sizeof(int));
    {
        // A/B switch for profiling: RT_B200_TRAVERSAL=simple selects the one-thread-per-ray kernels
        const char* e = getenv("RT_B200_TRAVERSAL");
        r->persistent = !(e && strcmp(e, "simple") == 0);
        // path-tracer schedule: params->schedule, overridable for A/B runs with RT_B200_PT_SCHEDULE
        r->useStreams = params->schedule != RT_SCHEDULE_WAVEFRONT;
        if ((e = getenv("RT_B200_PT_SCHEDULE")) != nullptr) r->useStreams = strcmp(e, "wavefront") != 0;
        const bool altAccel = is_alt_kind(scene->d.kind);
        if ((e = getenv("RT_B200_STREAM_KERNEL")) != nullptr && atoi(e) > 0) r->streamKernel = atoi(e);
        if (altAccel) r->streamKernel = 0; // k_pt_streams_alt: versions 5 / 8 are state machines over the BVH layout
        if ((e = getenv("RT_B200_STREAM_LPT")) != nullptr) r->streamLpt = atoi(e) != 0;
        int occ = 0;
        cudaError_t oe;
        if ((e = getenv("RT_B200_STREAM_MEASURED_LPT")) != nullptr) r->streamMeasuredLpt = atoi(e) != 0;
        if ((e = getenv("RT_B200_ORDERED_FRAMES")) != nullptr) r->orderedFrames = atoi(e) != 0;
        if ((e = getenv("RT_B200_IMAGE_BUDGET_MB")) != nullptr && atoll(e) > 0) r->imageBudgetBytes = (size_t)atoll(e) << 20;
        if (r->streamKernel != 5 && r->streamKernel != 8 && !altAccel) r->streamKernel = 8;
        if (r->streamKernel == 5 || r->streamKernel == 8)
        {
            if ((e = getenv("RT_B200_STREAM_MINB")) != nullptr && atoi(e) > 0) r->streamMinB = atoi(e);
            if ((e = getenv("RT_B200_STREAM_KEEPSHIFT")) != nullptr && atoi(e) > 0) r->streamKeepShift = atoi(e);
            if ((e = getenv("RT_B200_STREAM_LANECAP")) != nullptr) r->streamLaneCap = atoi(e) != 0;
            if ((e = getenv("RT_B200_STREAM_FASTNODE")) != nullptr) r->streamFastNode = atoi(e) != 0;
            if (r->streamKernel == 5 && !getenv("RT_B200_STREAM_KEEPSHIFT")) r->streamKeepShift = 2; // version 5's tuned value
            const bool perPixel = params->seed_mode == RT_SEED_PER_PIXEL;
            r->streamSmemSlots = streams8_slots(scene, perPixel);
            oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, r->streamKernel == 8 ? streams8_kernel(scene->d.kind == RT_SCENE_TLAS, r->streamMinB, perPixel, r->streamSmemSlots)
                                                                                           : streams5_kernel(scene->d.kind == RT_SCENE_TLAS, r->streamMinB, perPixel), 128, 0);
            if (r->streamKernel == 8 && !perPixel && !getenv("RT_B200_STREAM_MINB") && !getenv("RT_B200_STREAM_CTAS"))
            {
                int occ8 = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ8, streams8_kernel(scene->d.kind == RT_SCENE_TLAS, 8, perPixel, r->streamSmemSlots), 128, 0) == cudaSuccess && occ8 > 0)
                    r->streamMinBAuto = true, r->streamCtasPerSm8 = occ8;
                else cudaGetLastError();
            }
        }
        else if (scene->d.kind == RT_SCENE_FLAT_KDTREE) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pt_streams_alt<ACCEL_KD>, 128, 0);
        else if (scene->d.kind == RT_SCENE_FLAT_GRID) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pt_streams_alt<ACCEL_GRID>, 128, 0);
        else if (scene->d.kind == RT_SCENE_TLAS_KDTREE) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pt_streams_alt<ACCEL_TLAS_KD>, 128, 0);
        else oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pt_streams_alt<ACCEL_TLAS_GRID>, 128, 0);
        if (oe == cudaSuccess && occ > 0) r->streamCtasPerSm = occ;
        if ((e = getenv("RT_B200_STREAM_CTAS")) != nullptr && atoi(e) > 0) r->streamCtasPerSm = atoi(e);
    }
    if (cudaMallocHost((void**)&r->hCount, 8 * sizeof(int)) != cudaSuccess) { set_error("pinned alloc failed"); return fail(RT_ERR_CUDA); }
    if ((st = ralloc(r, &r->dPixels, px * 4)) != RT_OK) return fail(st);

    if (params->integrator == RT_INTEGRATOR_WHITTED)
    {
        WhState& w = r->wh;
        size_t cap = 4 * px;
        if (const char* e = getenv("RT_B200_WHITTED_QUEUE_PER_PIXEL")) { if (atof(e) > 0) cap = (size_t)(atof(e) * px) + 1; } // tests: force the overflow path
        if ((st = whitted_alloc_queues(r, cap)) != RT_OK) return fail(st);
        w.count = r->dCount, w.counters = r->dCounters;
        if ((st = ralloc(r, &r->dCam, sizeof(DCamera))) != RT_OK) return fail(st);
        { const char* e = getenv("RT_B200_WHITTED_GRAPH"); if (e) r->whGraphEnabled = atoi(e) != 0; }
        w.W = params->width, w.H = params->height, w.depthLimit = params->depth_limit, w.eps = params->epsilon;
    }
    *out = r;
    return RT_OK;
}

void rt_renderer_destroy(rt_renderer* r)
{
    if (!r) return;
    cudaSetDevice(r->scene->device);
    if (r->stream) cudaStreamSynchronize(r->stream);
    if (r->importedAccum) cudaIpcCloseMemHandle(r->importedAccum);
    for (void* p : r->allocations) cudaFree(p);
    for (cudaEvent_t e : r->evPool) cudaEventDestroy(e);
    if (r->hCount) cudaFreeHost(r->hCount);
    if (r->whGraphExec) cudaGraphExecDestroy(r->whGraphExec);
    if (r->ownStream) cudaStreamDestroy(r->ownStream);
    rt_scene* scene = r->scene;
    delete r;
    if (scene->renderers.fetch_sub(1) == 1 && scene->destroy_requested.load()) rt_scene_destroy(scene);
}

rt_status rt_renderer_set_stream(rt_renderer* r, void* stream)
{
    if (!r) return RT_ERR_INVALID;
    r->stream = stream ? (cudaStream_t)stream : r->ownStream;
    return RT_OK;
}

// Renderer::passes (the UI's "spp" slider, 3. PathTracer/renderer.cpp:182): samples per pixel per Tick
rt_status rt_renderer_set_passes(rt_renderer* r, int passes)
{
    if (!r || passes < 1) { set_error("rt_renderer_set_passes: bad argument"); return RT_ERR_INVALID; }
    if (passes > 8) { set_error("rt_renderer_set_passes: passes > 8 (the reference's slider stops at 4)"); return RT_ERR_UNSUPPORTED; }
    if (passes != r->passes) r->aheadValid = false; // frames rendered ahead were rendered with the old value
    r->passes = passes;
    return RT_OK;
}

rt_status rt_renderer_set_accumulator(rt_renderer* r, void* d_accumulator)
{
    if (!r) return RT_ERR_INVALID;
    r->accum = d_accumulator ? (float4*)d_accumulator : r->ownAccum;
    return RT_OK;
}

rt_status rt_renderer_set_camera(rt_renderer* r, const rt_camera* cam)
{
    if (!r || !cam) return RT_ERR_INVALID;
    const DCamera c = make_camera(*cam, r->params.width, r->params.height);
    if (memcmp(&c, &r->cam, sizeof c) != 0) r->tileOrderValid = false, r->aheadValid = false; // tile costs and frames rendered ahead are per view
    r->cam = c;
    return RT_OK;
}

rt_status rt_renderer_clear(rt_renderer* r)
{
    if (!r) return RT_ERR_INVALID;
    RT_CUDA(cudaSetDevice(r->scene->device));
    const rt_render_params& P = r->params;
    const int all = (P.width / 16) * (P.height / 16), mine = num_tiles(P);
    if (P.integrator == RT_INTEGRATOR_PATH && mine < all && mine > 0)
    {
        // a tile shard clears its own tiles only: the accumulator may be shared with the other shards (peer-mapped)
        k_clear_tiles<<<mine < r->sms * 8 ? mine : r->sms * 8, 256, 0, r->stream>>>(r->accum, P.width, P.width / 16, P.tile_begin, P.tile_step > 0 ? P.tile_step : 1, mine);
        r->launches++;
        RT_CUDA(cudaGetLastError());
        return RT_OK;
    }
    RT_CUDA(cudaMemsetAsync(r->accum, 0, (size_t)P.width * P.height * 16, r->stream));
    return RT_OK;
}

rt_status rt_renderer_export_accumulator(rt_renderer* r, void* handle_out)
{
    if (!r || !handle_out) { set_error("rt_renderer_export_accumulator: null argument"); return RT_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == RT_IPC_HANDLE_BYTES, "IPC handle size");
    if (r->accum != r->ownAccum) { set_error("rt_renderer_export_accumulator: only the renderer's own accumulator can be exported"); return RT_ERR_INVALID; }
    RT_CUDA(cudaSetDevice(r->scene->device));
    cudaIpcMemHandle_t h;
    RT_CUDA(cudaIpcGetMemHandle(&h, r->ownAccum));
    memcpy(handle_out, &h, sizeof h);
    return RT_OK;
}

rt_status rt_renderer_import_accumulator(rt_renderer* r, const void* handle)
{
    if (!r || !handle) { set_error("rt_renderer_import_accumulator: null argument"); return RT_ERR_INVALID; }
    RT_CUDA(cudaSetDevice(r->scene->device));
    RT_CUDA(cudaStreamSynchronize(r->stream));
    if (r->importedAccum) { cudaIpcCloseMemHandle(r->importedAccum); r->importedAccum = nullptr; r->accum = r->ownAccum; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    RT_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    r->importedAccum = p, r->accum = (float4*)p;
    r->aheadValid = false;
    return RT_OK;
}

static rt_status pt_ensure_slots(rt_renderer* r, int slots)
{
    if (slots <= r->ptSlotsAllocated) return RT_OK;
    // (re)allocate; old buffers stay in the allocation list until destroy (growth happens at most a few times)
    PTState& p = r->pt;
    rt_status st;
    const size_t S = (size_t)slots;
    if ((st = ralloc(r, &p.rayO, S * 16)) != RT_OK) return st;
    if ((st = ralloc(r, &p.rayD, S * 16)) != RT_OK) return st;
    if ((st = ralloc(r, &p.hit, S * 16)) != RT_OK) return st;
    if ((st = ralloc(r, &p.hitTri, S * 4)) != RT_OK) return st;
    if ((st = ralloc(r, &p.seed, S * 4)) != RT_OK) return st;
    if ((st = ralloc(r, &p.pix, S * 4)) != RT_OK) return st;
    const size_t levels = r->params.depth_limit > 0 ? r->params.depth_limit : 1;
    if ((st = ralloc(r, &p.weights, S * 16 * levels)) != RT_OK) return st;
    if ((st = ralloc(r, &p.active[0], S * 4)) != RT_OK) return st;
    if ((st = ralloc(r, &p.active[1], S * 4)) != RT_OK) return st;
    if (!p.history && (st = ralloc(r, &p.history, (size_t)(256 * 8 * 65 + 2) * 4)) != RT_OK) return st;
    r->ptSlotsAllocated = slots;
    return RT_OK;
}

// tile order for the stream pool: pilot cost per tile, sorted descending on the host (one small
// round trip per render call; the order depends on scene, camera and tile range only, so it is cached)
static rt_status pt_tile_order(rt_renderer* r, const PTState& p)
{
    const int n = p.nTiles;
    if (r->tileOrderValid && r->tileOrderCount == n)
    {
        if (r->tileOrderSource == 1 && r->tileClockRecorded && r->streamMeasuredLpt)
        {
            // the previous launch of this view timed every stream: re-sort by measured tile cost, once
            std::vector<unsigned long long> clk(n);
            RT_CUDA(cudaMemcpyAsync(clk.data(), r->dTileClock, (size_t)n * 8, cudaMemcpyDeviceToHost, r->stream));
            RT_CUDA(cudaStreamSynchronize(r->stream));
            std::vector<int> order(n);
            for (int i = 0; i < n; i++) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return clk[a] > clk[b]; });
            if (getenv("RT_B200_DEBUG"))
            {
                double sum = 0;
                for (int i = 0; i < n; i++) sum += (double)clk[i];
                fprintf(stderr, "[rt_b200] measured tile cost (cycles per stream, launch of %d frames): max %.0f  p99 %.0f  median %.0f  mean %.0f\n",
                    r->lastStreamFrames, (double)clk[order[0]] / r->lastStreamFrames, (double)clk[order[n / 100]] / r->lastStreamFrames,
                    (double)clk[order[n / 2]] / r->lastStreamFrames, sum / n / r->lastStreamFrames);
            }
            RT_CUDA(cudaMemcpyAsync(r->dTileOrder, order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, r->stream));
            RT_CUDA(cudaStreamSynchronize(r->stream));
            r->tileOrderSource = 2;
        }
        return RT_OK;
    }
    if (r->tileOrderCapacity < n)
    {
        rt_status st;
        if ((st = ralloc(r, &r->dTileOrder, (size_t)n * 4)) != RT_OK) return st;
        if ((st = ralloc(r, &r->dTileCost, (size_t)n * 4)) != RT_OK) return st;
        if ((st = ralloc(r, &r->dTileClock, (size_t)n * 8)) != RT_OK) return st;
        r->tileOrderCapacity = n;
    }
    RT_CUDA(cudaMemsetAsync(r->dTileCost, 0, (size_t)n * 4, r->stream));
    r->prof_begin();
    RT_FOR_ACCEL(r->scene->d.kind, (k_pt_pilot<A><<<r->sms * 4, 128, 0, r->stream>>>(p, r->scene->d, r->cam, r->dTileCost)));
    r->prof_end(RT_STAGE_GENERATE);
    std::vector<unsigned int> cost(n);
    RT_CUDA(cudaMemcpyAsync(cost.data(), r->dTileCost, (size_t)n * 4, cudaMemcpyDeviceToHost, r->stream));
    RT_CUDA(cudaStreamSynchronize(r->stream));
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    RT_CUDA(cudaMemcpyAsync(r->dTileOrder, order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, r->stream));
    RT_CUDA(cudaStreamSynchronize(r->stream)); // `order` is pageable host memory
    r->tileOrderValid = true, r->tileOrderCount = n;
    r->tileOrderSource = 1, r->tileClockRecorded = false;
    return RT_OK;
}

static rt_status render_pt_streams(rt_renderer* r, int first_spp, int count, int stride, float4* frameBuf = nullptr, bool compactImages = false)
{
    // the stream kernels keep the frame of a launch in 18 bits of a lane's sample counter (rt_streams8.cuh PIX_FRAME_MASK)
    constexpr int MAX_FRAMES_PER_LAUNCH = 1 << 17;
    if (count > MAX_FRAMES_PER_LAUNCH)
    {
        if (frameBuf) { set_error("rt_renderer_render: more than 131 072 frames in one launch with sample images"); return RT_ERR_UNSUPPORTED; }
        for (int done = 0; done < count; done += MAX_FRAMES_PER_LAUNCH)
        {
            const int frames = count - done < MAX_FRAMES_PER_LAUNCH ? count - done : MAX_FRAMES_PER_LAUNCH;
            const rt_status st = render_pt_streams(r, first_spp + done * stride, frames, stride);
            if (st != RT_OK) return st;
        }
        return RT_OK;
    }
    const rt_render_params& P = r->params;
    const int nTiles = num_tiles(P);
    PTState p = {};
    p.counters = r->dCounters, p.accum = r->accum, p.frameBuf = frameBuf;
    p.frameCompact = frameBuf && compactImages ? 1 : 0;
    p.nTiles = nTiles, p.tilesX = P.width / 16, p.tileBegin = P.tile_begin, p.tileStep = P.tile_step > 0 ? P.tile_step : 1;
    p.invTileStep = 1.0f / (float)p.tileStep;
    p.W = P.width, p.H = P.height, p.depthLimit = P.depth_limit, p.seedMode = P.seed_mode, p.eps = P.epsilon, p.passes = r->passes;
    p.stride = stride, p.firstSpp = first_spp;
    // all frames of the call form one pool of nTiles x count streams (x 256 with one stream per pixel; int range checked by the caller)
    const bool perPixel = P.seed_mode == RT_SEED_PER_PIXEL;
    p.slots = nTiles * count * (perPixel ? 256 : 1);
    RT_CUDA(cudaMemsetAsync(r->dCount + 6, 0, sizeof(int), r->stream));
    const int* order = nullptr;
    if (r->streamLpt && !perPixel) // per-pixel streams are one path long: nothing to balance
    {
        rt_status st = pt_tile_order(r, p);
        if (st != RT_OK) return st;
        order = r->dTileOrder;
    }
    // time the streams of this launch only while the tile order still comes from the pilot
    // (the BVH kernel only: on the KD-tree / grid kernel the measured order was no better than the pilot's, profiles/r1_kdtree_grid_*)
    unsigned long long* clk = (order && r->tileOrderSource == 1 && r->streamMeasuredLpt && r->streamKernel != 0) ? r->dTileClock : nullptr;
    if (clk) RT_CUDA(cudaMemsetAsync(clk, 0, (size_t)nTiles * 8, r->stream));
    r->prof_begin();
    if (r->streamKernel != 0)
    {
        // small jobs: spread the streams over all resident warps instead of filling the first warps completely
        int minb = r->streamMinB, ctasPerSm = r->streamCtasPerSm;
        if (r->streamMinBAuto && (r->scene->d.kind == RT_SCENE_TLAS || (long long)p.slots >= 8ll * r->sms * r->streamCtasPerSm8 * 128)) minb = 8, ctasPerSm = r->streamCtasPerSm8;
        const long long warps = (long long)r->sms * ctasPerSm * 4;
        long long perWarp = ((long long)p.slots + warps - 1) / warps;
        if (perWarp < 1) perWarp = 1;
        if (perWarp > 32 || !r->streamLaneCap) perWarp = 32;
        if (const char* e = getenv("RT_B200_STREAM_FORCE_LANES")) { const int v = atoi(e); if (v >= 1 && v <= 32) perWarp = v; } // experiments: streams per warp
        const unsigned laneMask = perWarp >= 32 ? 0xffffffffu : ((1u << perWarp) - 1);
        // (a refit with a TLAS rebuild can deepen the scene after the renderer chose its shared-memory stack: fall back to the local one)
        const int smemSlots = r->streamSmemSlots > r->scene->stack_entries ? r->streamSmemSlots : 0;
        const Streams5Fn fn = r->streamKernel == 8 ? streams8_kernel(r->scene->d.kind == RT_SCENE_TLAS, minb, perPixel, smemSlots)
                                                   : streams5_kernel(r->scene->d.kind == RT_SCENE_TLAS, minb, perPixel);
        const int ks = r->streamKeepShift | (r->streamKernel == 8 && r->streamFastNode ? 256 : 0);
        fn<<<r->sms * ctasPerSm, 128, 0, r->stream>>>(p, r->scene->d, r->cam, order, count, r->dCount + 6, clk, ks, laneMask);
    }
    else if (r->scene->d.kind == RT_SCENE_FLAT_KDTREE) k_pt_streams_alt<ACCEL_KD><<<r->sms * r->streamCtasPerSm, 128, 0, r->stream>>>(p, r->scene->d, r->cam, order, count, r->dCount + 6);
    else if (r->scene->d.kind == RT_SCENE_FLAT_GRID) k_pt_streams_alt<ACCEL_GRID><<<r->sms * r->streamCtasPerSm, 128, 0, r->stream>>>(p, r->scene->d, r->cam, order, count, r->dCount + 6);
    else if (r->scene->d.kind == RT_SCENE_TLAS_KDTREE) k_pt_streams_alt<ACCEL_TLAS_KD><<<r->sms * r->streamCtasPerSm, 128, 0, r->stream>>>(p, r->scene->d, r->cam, order, count, r->dCount + 6);
    else k_pt_streams_alt<ACCEL_TLAS_GRID><<<r->sms * r->streamCtasPerSm, 128, 0, r->stream>>>(p, r->scene->d, r->cam, order, count, r->dCount + 6);
    r->prof_end(RT_STAGE_EXTEND);
    if (clk) r->tileClockRecorded = true, r->lastStreamFrames = count;
    r->paths += (uint64_t)nTiles * count * 256 * r->passes;
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

// Grows the per-sample image buffer to `images` W x H float4 images; on an allocation failure the request is halved until one
// frame's images fit (the caller then renders fewer frames per launch).  Returns the number of images available, 0 = none.
static size_t ensure_images(rt_renderer* r, size_t images, size_t atLeast, size_t pixelsPerImage = 0)
{
    const size_t px = pixelsPerImage ? pixelsPerImage : (size_t)r->params.width * r->params.height;
    if (images * px <= r->imagesCapacity) return images;
    if (r->dImages)
    {
        cudaStreamSynchronize(r->stream);
        for (size_t i = 0; i < r->allocations.size(); i++)
            if (r->allocations[i] == r->dImages) { r->allocations.erase(r->allocations.begin() + i); break; }
        cudaFree(r->dImages);
        r->dImages = nullptr, r->imagesCapacity = 0;
    }
    while (true)
    {
        void* p = nullptr;
        if (cudaMalloc(&p, images * px * 16) == cudaSuccess)
        {
            r->dImages = (float4*)p, r->imagesCapacity = images * px;
            r->allocations.push_back(p);
            return images;
        }
        cudaGetLastError();
        if (images <= atLeast) return 0;
        images = images / 2 < atLeast ? atLeast : images / 2;
    }
}

// A job of `count` frames with the accumulator updated the way the reference's Tick sequence updates it (renderer.cpp:124:
// accumulator[pixel] += sample, frame after frame, pass after pass): the frames of a launch write their samples to separate
// images, k_sum_frames adds the images in order.  The result does not depend on how the launch was scheduled: it is bit-identical
// to `count` single-frame calls and to the reference (tests/test_gpu_parity.py).  Launches hold as many frames as fit the image
// budget (1080p: 33 MB per image, 64 frames = 2.1 GB; the sum pass reads them once: < 1 ms of HBM time).
static rt_status render_pt_streams_ordered(rt_renderer* r, int first_spp, int count, int stride)
{
    const rt_render_params& P = r->params;
    const bool kernelWritesImages = r->streamKernel == 8 || r->streamKernel == 0 || r->passes == 1; // version 5 indexes images by frame only
    if (!r->orderedFrames || !kernelWritesImages) return render_pt_streams(r, first_spp, count, stride);
    const int nTiles = num_tiles(P);
    // version 8 writes compact images (the job's own tiles only): a tile shard of N ranks holds N times as many frames per launch
    const bool compact = r->streamKernel == 8;
    const size_t px = compact ? (size_t)nTiles * 256 : (size_t)P.width * P.height;
    const size_t perFrame = (size_t)r->passes;
    size_t framesPerLaunch = r->imageBudgetBytes / (px * 16 * perFrame);
    if (framesPerLaunch < 1) framesPerLaunch = 1;
    if (framesPerLaunch > (size_t)count) framesPerLaunch = (size_t)count;
    if (framesPerLaunch > (1u << 17)) framesPerLaunch = 1u << 17; // render_pt_streams: MAX_FRAMES_PER_LAUNCH
    const size_t got = ensure_images(r, framesPerLaunch * perFrame, perFrame, px);
    if (got == 0) { set_error("rt_renderer_render: no device memory for one frame's sample images"); return RT_ERR_CUDA; }
    framesPerLaunch = got / perFrame;
    for (int done = 0; done < count; done += (int)framesPerLaunch)
    {
        const int frames = count - done < (int)framesPerLaunch ? count - done : (int)framesPerLaunch;
        rt_status st = render_pt_streams(r, first_spp + done * stride, frames, stride, r->dImages, compact);
        if (st != RT_OK) return st;
        r->prof_begin();
        k_sum_frames<<<nTiles < r->sms * 8 ? nTiles : r->sms * 8, 256, 0, r->stream>>>(r->accum, r->dImages, frames * (int)perFrame, P.width, P.height, P.width / 16,
            P.tile_begin, P.tile_step > 0 ? P.tile_step : 1, nTiles, compact ? 1 : 0);
        r->prof_end(RT_STAGE_ACCUMULATE);
        RT_CUDA(cudaGetLastError());
    }
    return RT_OK;
}

static rt_status render_pt(rt_renderer* r, int first_spp, int count, int stride)
{
    const rt_render_params& P = r->params;
    const int nTiles = num_tiles(P);
    if (nTiles == 0 || count <= 0) return RT_OK;
    const long long streamsPerFrame = (long long)nTiles * (P.seed_mode == RT_SEED_PER_PIXEL ? 256 : 1);
    if (r->useStreams && (P.seed_mode == RT_SEED_REFERENCE_TILE || r->streamKernel != 0) && P.depth_limit <= STREAM_MAX_DEPTH
        && streamsPerFrame * count < (1ll << 30))
    {
        const int L = P.lookahead_frames;
        // (with passes > 1 a frame image would hold the SUM of a pixel's samples, which the accumulator then receives in
        // one add instead of `passes` adds: not the Tick sequence bit for bit, so look-ahead serves passes == 1 only)
        if (count == 1 && L > 1 && r->streamKernel != 0 && r->passes == 1 && streamsPerFrame * L < (1ll << 30))
        {
            // one Tick per call: serve the frame from the images rendered ahead, rendering L more when it is not there
            const size_t px = (size_t)P.width * P.height;
            const bool hit = r->aheadValid && r->aheadStride == stride && first_spp >= r->aheadBase &&
                             (first_spp - r->aheadBase) % stride == 0 && (first_spp - r->aheadBase) / stride < r->aheadReady;
            if (!hit)
            {
                if (r->aheadCapacity < L)
                {
                    rt_status st = ralloc(r, &r->dFrameBuf, px * 16 * (size_t)L);
                    if (st != RT_OK) return st;
                    // pixels outside the rendered tiles (1080 % 16 rows, tile ranges of other shards) stay zero
                    RT_CUDA(cudaMemsetAsync(r->dFrameBuf, 0, px * 16 * (size_t)L, r->stream));
                    r->aheadCapacity = L;
                }
                rt_status st = render_pt_streams(r, first_spp, L, stride, r->dFrameBuf);
                if (st != RT_OK) return st;
                r->aheadValid = true, r->aheadBase = first_spp, r->aheadStride = stride, r->aheadReady = L;
            }
            const int k = (first_spp - r->aheadBase) / stride;
            // (the shard's own tiles only: the accumulator may be shared with other tile shards)
            r->prof_begin();
            k_sum_frames<<<nTiles < r->sms * 8 ? nTiles : r->sms * 8, 256, 0, r->stream>>>(r->accum, r->dFrameBuf + (size_t)k * px, 1, P.width, P.height, P.width / 16,
                P.tile_begin, P.tile_step > 0 ? P.tile_step : 1, nTiles, 0);
            r->prof_end(RT_STAGE_ACCUMULATE);
            RT_CUDA(cudaGetLastError());
            return RT_OK;
        }
        return render_pt_streams_ordered(r, first_spp, count, stride);
    }
    // wavefront: slots per frame = tiles (reference RNG: a slot walks its tile) or pixels (one stream per pixel)
    const bool wfPerPixel = P.seed_mode == RT_SEED_PER_PIXEL;
    const int slotsPerFrame = nTiles * (wfPerPixel ? 256 : 1);
    int inFlight = P.max_frames_in_flight > 0 ? P.max_frames_in_flight : (wfPerPixel ? 8 << 20 : 1 << 20) / slotsPerFrame;
    if (inFlight < 1) inFlight = 1;
    if (inFlight > count) inFlight = count;
    // ordered accumulation as for the stream schedule: the frames in flight write their samples to (frame, pass) images
    float4* images = nullptr;
    if (r->orderedFrames)
    {
        const size_t px = (size_t)P.width * P.height;
        size_t fit = r->imageBudgetBytes / (px * 16 * (size_t)r->passes);
        if (fit < 1) fit = 1;
        if ((size_t)inFlight > fit) inFlight = (int)fit;
        const size_t got = ensure_images(r, (size_t)inFlight * r->passes, (size_t)r->passes);
        if (got == 0) { set_error("rt_renderer_render: no device memory for one frame's sample images"); return RT_ERR_CUDA; }
        inFlight = (int)(got / r->passes);
        images = r->dImages;
    }
    rt_status st = pt_ensure_slots(r, slotsPerFrame * inFlight);
    if (st != RT_OK) return st;
    PTState& p = r->pt;
    p.count = r->dCount, p.counters = r->dCounters, p.accum = r->accum, p.frameBuf = images;
    p.nTiles = nTiles, p.tilesX = P.width / 16, p.tileBegin = P.tile_begin, p.tileStep = P.tile_step > 0 ? P.tile_step : 1;
    p.W = P.width, p.H = P.height, p.depthLimit = P.depth_limit, p.seedMode = P.seed_mode, p.eps = P.epsilon, p.passes = r->passes;
    p.stride = stride;
    const int grid = r->sms * 8;
    const int kind = r->scene->d.kind;
    // every path makes at most depth_limit + 1 FindNearest queries, every slot 256 paths
    const int maxIters = (wfPerPixel ? 1 : 256) * r->passes * (P.depth_limit + 1) + 1;
    for (int done = 0; done < count; done += inFlight)
    {
        const int frames = count - done < inFlight ? count - done : inFlight;
        p.slots = slotsPerFrame * frames;
        p.firstSpp = first_spp + done * stride;
        r->prof_begin();
        k_pt_generate<<<r->sms * 4, 256, 0, r->stream>>>(p, r->cam);
        r->prof_end(RT_STAGE_GENERATE);
        int cur = 0;
        r->ptIterations = 0;
        for (int it = 0; it < maxIters; it++)
        {
            p.iteration = it;
            r->ptIterations = it + 1;
            r->prof_begin();
            if (r->persistent) { RT_FOR_ACCEL(kind, (k_pt_extend_persistent<A><<<grid, 128, 0, r->stream>>>(p, r->scene->d, cur))); }
            else { RT_FOR_ACCEL(kind, (k_pt_extend<A><<<grid, 128, 0, r->stream>>>(p, r->scene->d, cur))); }
            r->prof_end(RT_STAGE_EXTEND);
            r->prof_begin();
            k_pt_shade<<<grid, 128, 0, r->stream>>>(p, r->scene->d, r->cam, cur);
            r->prof_end(RT_STAGE_SHADE);
            cur ^= 1;
            if ((it & 31) == 31)
            {
                RT_CUDA(cudaMemcpyAsync(r->hCount, p.count + cur, sizeof(int), cudaMemcpyDeviceToHost, r->stream));
                RT_CUDA(cudaStreamSynchronize(r->stream));
                if (r->hCount[0] == 0) break;
            }
        }
        if (images)
        {
            r->prof_begin();
            k_sum_frames<<<nTiles < r->sms * 8 ? nTiles : r->sms * 8, 256, 0, r->stream>>>(r->accum, images, frames * r->passes, P.width, P.height, P.width / 16,
                P.tile_begin, P.tile_step > 0 ? P.tile_step : 1, nTiles, 0);
            r->prof_end(RT_STAGE_ACCUMULATE);
        }
        r->paths += (uint64_t)nTiles * frames * 256 * r->passes;
    }
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

// the launches of one Whitted frame (Renderer::Tick, 2. WhittedStyle/renderer.cpp:131-157)
static void whitted_frame_launches(rt_renderer* r)
{
    const rt_render_params& P = r->params;
    WhState& w = r->wh;
    const size_t px = (size_t)P.width * P.height;
    // Tick overwrites the accumulator every frame (renderer.cpp:155)
    cudaMemsetAsync(r->accum, 0, px * 16, r->stream);
    r->prof_begin();
    k_wh_generate<<<r->sms * 4, 256, 0, r->stream>>>(w, r->dCam);
    r->prof_end(RT_STAGE_GENERATE);
    const int grid = r->sms * 8;
    const int kind = r->scene->d.kind;
    int cur = 0;
    for (int depth = 0; depth <= P.depth_limit; depth++)
    {
        r->prof_begin();
        if (r->persistent) { RT_FOR_ACCEL(kind, (k_wh_extend_persistent<A><<<grid, 128, 0, r->stream>>>(w, r->scene->d, cur))); }
        else { RT_FOR_ACCEL(kind, (k_wh_extend<A><<<grid, 128, 0, r->stream>>>(w, r->scene->d, cur))); }
        r->prof_end(RT_STAGE_EXTEND);
        r->prof_begin();
        k_wh_shade<<<grid, 128, 0, r->stream>>>(w, r->scene->d, cur);
        r->prof_end(RT_STAGE_SHADE);
        r->prof_begin();
        if (r->persistent) { RT_FOR_ACCEL(kind, (k_wh_connect_persistent<A><<<grid, 128, 0, r->stream>>>(w, r->scene->d))); }
        else { RT_FOR_ACCEL(kind, (k_wh_connect<A><<<grid, 128, 0, r->stream>>>(w, r->scene->d))); }
        r->prof_end(RT_STAGE_CONNECT);
        cur ^= 1;
    }
}

// A Whitted frame is 2 + 3 * (depthLimit + 1) short launches (20 at depth 5): at 640x360 the frame is bound by
// launch latency, so the sequence is captured once into a CUDA graph and replayed; only the camera (device
// memory) changes between frames.  The graph is rebuilt when the accumulator pointer changes.
static rt_status render_whitted(rt_renderer* r)
{
    const rt_render_params& P = r->params;
    r->wh.accum = r->accum;
    // pageable source: the runtime stages the 56 bytes before returning, so r->cam may change right after
    RT_CUDA(cudaMemcpyAsync(r->dCam, &r->cam, sizeof(DCamera), cudaMemcpyHostToDevice, r->stream));
    const bool useGraph = r->whGraphEnabled && !r->profiling;
    if (useGraph)
    {
        if (r->whGraphExec && r->whGraphAccum != r->accum)
        {
            cudaGraphExecDestroy(r->whGraphExec);
            r->whGraphExec = nullptr;
        }
        if (!r->whGraphExec)
        {
            cudaGraph_t g = nullptr;
            RT_CUDA(cudaStreamBeginCapture(r->stream, cudaStreamCaptureModeThreadLocal));
            const uint64_t before = r->launches;
            whitted_frame_launches(r);
            r->whGraphLaunches = (int)(r->launches - before);
            r->launches = before;
            RT_CUDA(cudaStreamEndCapture(r->stream, &g));
            RT_CUDA(cudaGraphInstantiate(&r->whGraphExec, g, 0));
            cudaGraphDestroy(g);
            r->whGraphAccum = r->accum;
        }
        RT_CUDA(cudaGraphLaunch(r->whGraphExec, r->stream));
        r->launches += r->whGraphLaunches;
    }
    else whitted_frame_launches(r);
    r->paths += (size_t)P.width * P.height;
    r->whLastCam = r->cam, r->whFramePending = true;
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

rt_status rt_renderer_render(rt_renderer* r, int first_spp, int count, int stride)
{
    if (!r) { set_error("rt_renderer_render: null renderer"); return RT_ERR_INVALID; }
    RT_CUDA(cudaSetDevice(r->scene->device));
    if (r->params.integrator == RT_INTEGRATOR_PATH) return render_pt(r, first_spp, count, stride > 0 ? stride : 1);
    return render_whitted(r);
}

rt_status rt_renderer_sync(rt_renderer* r)
{
    if (!r) return RT_ERR_INVALID;
    RT_CUDA(cudaSetDevice(r->scene->device));
    RT_CUDA(cudaStreamSynchronize(r->stream));
    if (r->params.integrator == RT_INTEGRATOR_WHITTED && r->whFramePending)
    {
        // a frame whose ray queues overflowed dropped rays: grow the queues and render it again (Tick overwrites the accumulator,
        // so the repeat is the frame), up to 2^depth_limit rays per pixel - the most a dielectric ray tree can hold at one depth
        for (int attempt = 0; attempt < 8; attempt++)
        {
            RT_CUDA(cudaMemcpy(r->hCount, r->dCount, 8 * sizeof(int), cudaMemcpyDeviceToHost));
            if (!r->hCount[3]) break;
            RT_CUDA(cudaMemset(r->dCount + 3, 0, sizeof(int))); // never sticky: later frames start clean
            const size_t px = (size_t)r->params.width * r->params.height;
            const size_t most = px << (r->params.depth_limit < 6 ? r->params.depth_limit : 6);
            if ((size_t)r->wh.capacity >= most || attempt == 7)
            {
                set_error("Whitted ray queue overflow: more rays alive at one depth than the queues can grow to");
                r->whFramePending = false;
                return RT_ERR_UNSUPPORTED;
            }
            rt_status st = whitted_alloc_queues(r, (size_t)r->wh.capacity * 2 < most ? (size_t)r->wh.capacity * 2 : most);
            if (st != RT_OK) return st;
            const DCamera now = r->cam;
            r->cam = r->whLastCam;
            st = render_whitted(r);
            r->cam = now;
            if (st != RT_OK) return st;
            r->paths -= px; // the repeat is not another frame
            RT_CUDA(cudaStreamSynchronize(r->stream));
        }
        r->whFramePending = false;
    }
    return RT_OK;
}

rt_status rt_renderer_read_accumulator(rt_renderer* r, float* host_rgba)
{
    if (!r || !host_rgba) return RT_ERR_INVALID;
    rt_status st = rt_renderer_sync(r);
    if (st != RT_OK) return st;
    RT_CUDA(cudaMemcpy(host_rgba, r->accum, (size_t)r->params.width * r->params.height * 16, cudaMemcpyDeviceToHost));
    return RT_OK;
}

rt_status rt_renderer_read_pixels(rt_renderer* r, float scale, uint32_t* host_rgb8)
{
    if (!r || !host_rgb8) return RT_ERR_INVALID;
    RT_CUDA(cudaSetDevice(r->scene->device));
    const int n = r->params.width * r->params.height;
    k_to_rgb8<<<r->sms * 4, 256, 0, r->stream>>>(r->accum, r->dPixels, n, scale);
    r->launches++;
    RT_CUDA(cudaMemcpyAsync(host_rgb8, r->dPixels, (size_t)n * 4, cudaMemcpyDeviceToHost, r->stream));
    RT_CUDA(cudaStreamSynchronize(r->stream));
    return RT_OK;
}

void* rt_renderer_device_accumulator(rt_renderer* r) { return r ? r->accum : nullptr; }

rt_status rt_renderer_get_counters(rt_renderer* r, rt_counters* out)
{
    if (!r || !out) return RT_ERR_INVALID;
    rt_status st = rt_renderer_sync(r);
    if (st != RT_OK) return st;
    unsigned long long c[4];
    RT_CUDA(cudaMemcpy(c, r->dCounters, sizeof c, cudaMemcpyDeviceToHost));
    out->extension_rays = c[0], out->shadow_rays = c[1], out->wavefront_iterations = c[2];
    out->paths = r->paths, out->kernel_launches = r->launches;
    return RT_OK;
}

rt_status rt_renderer_set_profiling(rt_renderer* r, int enabled)
{
    if (!r) return RT_ERR_INVALID;
    rt_status st = rt_renderer_sync(r);
    if (st != RT_OK) return st;
    r->profiling = enabled != 0;
    r->spans.clear(), r->evUsed = 0;
    return RT_OK;
}

rt_status rt_renderer_get_stage_times(rt_renderer* r, rt_stage_times* out)
{
    if (!r || !out) return RT_ERR_INVALID;
    rt_status st = rt_renderer_sync(r);
    if (st != RT_OK) return st;
    memset(out, 0, sizeof *out);
    for (const rt_renderer::Span& sp : r->spans)
    {
        float ms = 0;
        RT_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
        out->ms[sp.stage] += ms, out->launches[sp.stage]++;
    }
    r->spans.clear(), r->evUsed = 0;
    return RT_OK;
}

rt_status rt_renderer_get_launch_spans(rt_renderer* r, int32_t* stage, float* ms, size_t capacity, size_t* n)
{
    if (!r || !stage || !ms || !n) return RT_ERR_INVALID;
    rt_status st = rt_renderer_sync(r);
    if (st != RT_OK) return st;
    size_t k = 0;
    for (const rt_renderer::Span& sp : r->spans)
    {
        if (k == capacity) break;
        RT_CUDA(cudaEventElapsedTime(&ms[k], sp.a, sp.b));
        stage[k++] = sp.stage;
    }
    *n = k;
    r->spans.clear(), r->evUsed = 0;
    return RT_OK;
}

rt_status rt_renderer_get_queue_history(rt_renderer* r, int32_t* out, size_t capacity, size_t* n)
{
    if (!r || !out || !n) return RT_ERR_INVALID;
    rt_status st = rt_renderer_sync(r);
    if (st != RT_OK) return st;
    size_t k = (size_t)r->ptIterations < capacity ? (size_t)r->ptIterations : capacity;
    if (k && r->pt.history) RT_CUDA(cudaMemcpy(out, r->pt.history, k * 4, cudaMemcpyDeviceToHost));
    *n = r->pt.history ? k : 0;
    return RT_OK;
}

rt_status rt_renderer_reset_counters(rt_renderer* r)
{
    if (!r) return RT_ERR_INVALID;
    rt_status st = rt_renderer_sync(r);
    if (st != RT_OK) return st;
    RT_CUDA(cudaMemset(r->dCounters, 0, 4 * sizeof(unsigned long long)));
    r->paths = 0, r->launches = 0;
    return RT_OK;
}

} // extern "C"
