// rt_streams8.cuh — the path tracer's stream kernel, version 8 (included by rt_render.cu after PTState / pt_surface).
//
// Same schedule as version 5 (one (tile, frame) RNG stream of the reference per lane, path state in registers, the warp runs the
// action most lanes wait for: NODE / LEAF / SHADE / MISS), rebuilt around what ncu showed on version 5
// (profiles/r2_v5_default_k_pt_streams5_ncu_full.txt, _source_hotspots.txt): 71 % of the issue slots busy at 14.7 of 32 lanes,
// 166 warp instructions per ray, 38 % of them in the interior-node step (96 SASS instructions), a third of the L1 load
// sectors going to the 64-entry local-memory stack.
//   * node step 96 -> ~60 instructions: the two slab tests use packed FADD2 / FMUL2 (rt_device.cuh slab_both, node layout
//     re-ordered so that box planes meet their ray constants pairwise); near / far by FMNMX instead of selects;
//   * the node stack is addressed through a running pointer: push = one predicated store, top-of-stack = one load, no index
//     arithmetic.  Its bottom slot holds CUR_END, so the pop that empties the stack needs no `sp == 0` test: the lane simply finds
//     CUR_END in `cur`.  Two placements of the same code: SMEM_SLOTS = 0, a 65-entry local-memory array (the DEFAULT: L1-resident,
//     and measured 1-4 % faster than the carve-out, profiles/r2_stream_kernel_sweeps.txt), or SMEM_SLOTS = 24 / 32, one column per
//     thread in shared memory (bank = lane: conflict-free; RT_B200_STREAM_SMEM_SLOTS, for trees that fit);
//   * `cur` is the whole traversal state (>= 0 interior node, leaf / instance references, CUR_* markers): the node loop
//     carries two registers (cur, sp) instead of packed bool flags;
//   * every sample is written to its own (frame, pass) image and the images are added to the accumulator in frame order by
//     k_sum_frames: a multi-frame call leaves the accumulator bit-identical to the reference's Tick sequence
//     (3. PathTracer/renderer.cpp:117-131) instead of float atomics in completion order;
//   * the glibc-exact expf / atan2f / acosf paths are out of line (rt_device.cuh beer_scale, sky_texel_exact_cold);
//   * ONE expansion of the NODE action for both ways into it (fast path and full vote): the kernel's hot code is larger than the
//     instruction cache and every KB shows (profiles/r2_code_footprint.txt: -3.5 % time on the flat scene; moving the NaN-exact
//     loop out of line or into a per-lane branch of the node step, also measured there, costs more than its 1 KB).
// Per lane the order of node visits, triangle tests, RNG draws and bounces is the reference's, as before.
#pragma once
#ifndef RT_S8_LIVE_BALLOT
#define RT_S8_LIVE_BALLOT 1
#endif
// 1: a TLAS ray's world-space slab constants are parked in local memory while it is inside an instance (-1 % time on the TLAS scenes,
// profiles/r2_stream_kernel_sweeps.txt block 6); 0: recomputed when the ray leaves the instance
#ifndef RT_S8_PARK_WORLD
#define RT_S8_PARK_WORLD 1
#endif

namespace rtb {

constexpr int CUR_END = (int)0x80000001u;   // popped from the bottom slot: the traversal of this ray is finished
constexpr int CUR_SHADE = (int)0x80000002u; // surface shading pending
constexpr int CUR_MISS = (int)0x80000003u;  // sky lookup pending
constexpr int CUR_DEAD = (int)0x80000004u;  // no stream
constexpr int CUR_EXIT = (int)0x80000005u;  // stack marker: leave the current instance (this kernel's SENTINEL)
// leaf / instance references are ~payload with payload < 0x7ffffffa (rt_scene_create bounds the instance count), i.e. > CUR_EXIT
// as unsigned numbers; a lane is in LEAF state iff (unsigned)cur > (unsigned)CUR_DEAD

// sample counter of a stream: pixel of the tile (0..256) | pass << 9 | frame of the launch << 13 | PIX_POOL_EMPTY (warp-uniform flag kept in
// every lane's counter: as a variable of its own it cost the TLAS instance of the kernel a local-memory load in front of every full vote)
constexpr int PIX_PASS_SHIFT = 9, PIX_FRAME_SHIFT = 13;
constexpr int PIX_POOL_EMPTY = (int)0x80000000u, PIX_FRAME_MASK = 0x3ffff; // frames of one launch < 2^18 (rt_renderer_render bounds them)

// One interior-node visit (bvh.cpp:242-257): both child boxes from one 64-byte record, near child first, left on ties, far
// child pushed only when hit, 1e30f compared with == like the reference.  STRIDE = distance between two stack slots of a lane.
template <bool EXACT, int STRIDE>
__device__ __forceinline__ void node_step8(const float4* __restrict__ nodes, const RaySlab& rs, const float ht, int*& sp, int& cur)
{
    const FatNode n = load_node(nodes, cur);
    const int top = sp[-STRIDE];
    float a1, a2;
    slab_both<EXACT>(rs, ht, n, a1, a2);
    const bool swp = a1 > a2;
    const float dn = fminf(a1, a2), df = fmaxf(a1, a2); // a1, a2 are never NaN (a NaN tmin fails the hit test: 1e30f)
    const int cn = swp ? n.right : n.left, cf = swp ? n.left : n.right;
    const bool miss = dn == 1e30f, both = !miss && df != 1e30f;
    if (both) *sp = cf;
    cur = miss ? top : cn;
    sp += both ? STRIDE : (miss ? -STRIDE : 0);
}

template <bool TLAS, int MINB, bool PERPIXEL, int SMEM_SLOTS>
__global__ void __launch_bounds__(128, MINB) k_pt_streams8(const PTState p, const DScene s, const DCamera cam,
    const int* __restrict__ tileOrder, const int frames, int* __restrict__ streamCounter, unsigned long long* __restrict__ tileCost, const int keepShiftAndFlags, const unsigned laneMask)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int total = p.slots;
    const float4* __restrict__ nodes = s.nodes;
    const float4* __restrict__ tris = s.tris;
    constexpr int STRIDE = SMEM_SLOTS > 0 ? 128 : 1;
    __shared__ int smemStack[SMEM_SLOTS > 0 ? SMEM_SLOTS * 128 : 1];
    // local-memory placement: 4 ints in front of the stack park the world-space reciprocal direction + degenerate-ray flag of a
    // TLAS ray while it is inside an instance (leaving the instance reloads them: no divisions, no re-test)
    constexpr bool PARK = RT_S8_PARK_WORLD && TLAS && SMEM_SLOTS == 0;
    __align__(16) int localStack[SMEM_SLOTS > 0 ? 1 : STACK_SIZE + 1 + (PARK ? 4 : 0)];
    int* const stackBase = SMEM_SLOTS > 0 ? smemStack + threadIdx.x : localStack + (PARK ? 4 : 0);
    stackBase[0] = CUR_END;
    int* sp = stackBase + STRIDE;
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
    int cur = CUR_DEAD, instObj = -1;
    float ht = 0, hu = 0, hv = 0;
    int hobj = -1, htri = -1;
    unsigned int rays = 0;
    const size_t imagePixels = (size_t)p.W * p.H;
#define TO (TLAS ? O : wO)
#define TD (TLAS ? D : wD)
    // NODE action: interior-node visits, repeated without a new vote while >= (1 - 2^-keepShift) of the lanes that entered are still
    // on interior nodes.  ONE expansion serves both ways into it (fast path and full vote).
    // The NaN-exact slab variant is chosen per ACTION, not per lane and box: if any lane of this action holds a degenerate ray
    // everyone takes the select-based min / max (same values for ordinary rays).
#define RT_S8_NODE_ACTION(NN)                                                                            \
    {                                                                                                    \
        const int keep = (NN) - ((NN) >> keepShift);                                                     \
        if (__any_sync(FULL, cur >= 0 && exact))                                                         \
        {                                                                                                \
            do                                                                                           \
            {                                                                                            \
                if (cur >= 0) node_step8<true, STRIDE>(nodes, rs, ht, sp, cur);                          \
            } while (__popc(__ballot_sync(FULL, cur >= 0)) >= keep);                                     \
        }                                                                                                \
        else                                                                                             \
        {                                                                                                \
            do                                                                                           \
            {                                                                                            \
                if (cur >= 0) node_step8<false, STRIDE>(nodes, rs, ht, sp, cur);                         \
            } while (__popc(__ballot_sync(FULL, cur >= 0)) >= keep);                                     \
        }                                                                                                \
        if (cur == CUR_END) cur = hobj == -1 ? CUR_MISS : CUR_SHADE;                                     \
    }

#if !RT_S8_LIVE_BALLOT
    int nLive = 32; // live lanes at the last full vote (warp-uniform)
#endif
    const bool fastNode = (keepShiftAndFlags & 256) != 0;
    const int keepShift = keepShiftAndFlags & 255;
    while (true)
    {
        const unsigned mNode = __ballot_sync(FULL, cur >= 0);
        const int nNodes = __popc(mNode);
        // fast path: interior-node lanes are at least half of the lanes that were live at the last full vote, so NODE wins any vote.
        // Skips the other three ballots, the counts and the refill test (dead lanes wait for the next full vote, which comes as
        // soon as NODE stops being the majority).
#if RT_S8_LIVE_BALLOT
        // (the live lanes are counted with a second ballot here rather than remembered from the last full vote: a register less, and
        // in the TLAS instance of the kernel ptxas kept that count in local memory - a load in front of every decision)
        const unsigned mLiveNow = __ballot_sync(FULL, cur != CUR_DEAD);
        const int nLive = __popc(mLiveNow);
#endif
        bool nodeAction = fastNode && 2 * nNodes >= nLive && nNodes > 0;
        bool start = false;
        if (!nodeAction)
        {
            const unsigned mLeaf = __ballot_sync(FULL, (unsigned)cur > (unsigned)CUR_DEAD);
            const unsigned mShade = __ballot_sync(FULL, cur == CUR_SHADE);
#if RT_S8_LIVE_BALLOT
            const unsigned mLive = mLiveNow, mMiss = mLive & ~(mNode | mLeaf | mShade); // a live lane is in exactly one of the four states
#else
            const unsigned mMiss = __ballot_sync(FULL, cur == CUR_MISS);
            const unsigned mLive = mNode | mLeaf | mShade | mMiss;
#endif
            // (per-pixel streams end after every path: refill in batches of >= 8 lanes so that refills do not alternate with actions)
            if ((~mLive & laneMask) != 0 && pix >= 0 && (!PERPIXEL || mLive == 0 || __popc(~mLive & laneMask) >= 8))
            {
                // refill the dead lanes from the stream pool: one atomic per warp.  laneMask caps the streams per warp
                // for small jobs (fewer streams than lanes): a chain runs faster the fewer neighbours it waits for
                const unsigned mDead = ~mLive & laneMask;
                const int nIdle = __popc(mDead);
                const int leader = __ffs(mDead) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(streamCounter, nIdle);
                base = __shfl_sync(FULL, base, leader);
                if (base + nIdle >= total) pix |= PIX_POOL_EMPTY;
                const int stream = base + __popc(mDead & ((1u << lane) - 1));
                if (cur == CUR_DEAD && ((laneMask >> lane) & 1) && stream < total)
                {
                    // RT_SEED_REFERENCE_TILE: a stream is a (tile, frame) pair and runs the tile's 256 pixels; RT_SEED_PER_PIXEL:
                    // a stream is ONE pixel of a (tile, frame) pair (32 consecutive streams = two pixel rows of one tile)
                    const int unit = PERPIXEL ? stream >> 8 : stream, px0 = PERPIXEL ? stream & 255 : 0;
                    const int k = unit / frames, frame = unit - k * frames;
                    const int tile = p.tileBegin + (tileOrder ? tileOrder[k] : k) * p.tileStep;
                    const int tx = tile % p.tilesX, ty = tile / p.tilesX;
                    const int x = tx * 16 + (px0 & 15), y = ty * 16 + (px0 >> 4);
                    seed = PERPIXEL ? pt_pixel_seed(p, x, y, p.firstSpp + frame * p.stride) : pt_seed(p, tile, p.firstSpp + frame * p.stride);
                    tileXY = (tx * 16) | ((ty * 16) << 16);
                    pix = (pix & PIX_POOL_EMPTY) | px0 | (frame << PIX_FRAME_SHIFT), depth = 0, inside = false;
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
                const int nN = nNodes, nL = __popc(mLeaf), nS = __popc(mShade), nM = __popc(mMiss);
#if !RT_S8_LIVE_BALLOT
                nLive = nN + nL + nS + nM;
#endif
                if (nN >= nL && nN >= nS && nN >= nM) nodeAction = true;
                else if (nL >= nS && nL >= nM)
                {
                    if ((unsigned)cur > (unsigned)CUR_DEAD)
                    {
                        const int payload = ~cur;
                        bool pop = true;
                        if (TLAS && cur == CUR_EXIT)
                        {
                            // blas_bvh.cpp:385-388
                            O = wO, D = wD;
                            if (PARK)
                            {
                                const int4 pk = *(const int4*)localStack;
                                rs = make_ray_slab(wO, f3(__int_as_float(pk.x), __int_as_float(pk.y), __int_as_float(pk.z))), exact = pk.w != 0;
                            }
                            else rs = make_ray_slab(wO, recip(wD)), exact = needs_exact_slab(wO, wD);
                        }
                        else if (TLAS && (payload & INSTANCE_BIT))
                        {
                            // TLAS leaf -> BLASBVH::Intersect (blas_bvh.cpp:376-389), SSE lane-sum order of TransformPosition_SSE /
                            // TransformVector_SSE (tmplmath.cpp:170-191)
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
                            *sp = CUR_EXIT, sp += STRIDE;
                            cur = meta.x;
                            pop = false;
                        }
                        else
                        {
                            // triangle leaf: bvh.cpp:232-241
                            int slot = payload;
                            while (true)
                            {
                                const float4* T = tris + 3 * (size_t)slot;
                                const float4 t0_ = __ldg(T), t1 = __ldg(T + 1), t2 = __ldg(T + 2);
                                const int tag = __float_as_int(t0_.w);
                                if (intersect_tri(TO, TD, f3(t0_.x, t0_.y, t0_.z), f3(t1.x, t1.y, t1.z), f3(t2.x, t2.y, t2.z), ht, hu, hv))
                                {
                                    htri = tag & ~LAST_BIT;
                                    const int own = TLAS ? instObj : s.flat_obj_idx; // BLAS' objIdx (blas_bvh.cpp:297) / Tri::objIdx (bvh.cpp:219)
                                    hobj = own >= 0 ? own : __float_as_int(t1.w);
                                }
                                if (tag & LAST_BIT) break;
                                slot++;
                            }
                        }
                        if (pop) sp -= STRIDE, cur = *sp;
                        if (cur == CUR_END) cur = hobj == -1 ? CUR_MISS : CUR_SHADE;
                    }
                }
                else
                {
                    // MISS (sky, renderer.cpp:54) or surface shading, whichever more lanes wait for; then ONE copy
                    // of "sample finished -> write it, next pixel" for the lanes whose path ended
                    const bool doMiss = nM >= nS;
                    bool fin = false;
                    float3 L = f3(0, 0, 0);
                    if (doMiss)
                    {
                        if (cur == CUR_MISS) L = sky_color(s, wD), fin = true;
                    }
                    else if (cur == CUR_SHADE)
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
                        int px = pix & 511, pass = (pix >> PIX_PASS_SHIFT) & 15; // `passes` consecutive samples per pixel (renderer.cpp:123)
                        const int frame = (pix >> PIX_FRAME_SHIFT) & PIX_FRAME_MASK;
                        const size_t pixel = (x0 + (px & 15)) + (size_t)(y0 + (px >> 4)) * p.W;
                        if (p.frameCompact)
                        {
                            // the sample's own image, compact layout: (k-th tile of the job) * 256 + pixel of the tile
                            const int k = __float2int_rn((float)((y0 >> 4) * p.tilesX + (x0 >> 4) - p.tileBegin) * p.invTileStep);
                            p.frameBuf[((size_t)(frame * p.passes + pass) * p.nTiles + k) * 256 + px] = make_float4(L.x, L.y, L.z, 0);
                        }
                        else if (p.frameBuf) p.frameBuf[(size_t)(frame * p.passes + pass) * imagePixels + pixel] = make_float4(L.x, L.y, L.z, 0); // W x H images (look-ahead)
                        else
                        {
                            float* a = (float*)(p.accum + pixel); // renderer.cpp:124, in completion order (callers that asked for no images)
                            atomicAdd(a + 0, L.x), atomicAdd(a + 1, L.y), atomicAdd(a + 2, L.z);
                        }
                        if (++pass == p.passes) pass = 0, px++;
                        pix = (pix & PIX_POOL_EMPTY) | px | (pass << PIX_PASS_SHIFT) | (frame << PIX_FRAME_SHIFT);
                        if (PERPIXEL ? pass != 0 : px < 256)
                        {
                            const float jy = random_float(seed), jx = random_float(seed);
                            wD = primary_dir(cam, (float)(x0 + (px & 15)) + jx, (float)(y0 + (px >> 4)) + jy);
                            wO = cam.pos, depth = 0, inside = false;
                            start = true;
                        }
                        else
                        {
                            cur = CUR_DEAD;
                            if (tileCost) atomicAdd(&tileCost[((y0 >> 4) * p.tilesX + (x0 >> 4) - p.tileBegin) / p.tileStep], (unsigned long long)((unsigned int)clock() - t0));
                        }
                    }
                }
            }
        }
        if (nodeAction) RT_S8_NODE_ACTION(nNodes);
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
            const float3 wrD = recip(wD);
            rs = make_ray_slab(wO, wrD), exact = needs_exact_slab(wO, wD);
            if (PARK) *(int4*)localStack = make_int4(__float_as_int(wrD.x), __float_as_int(wrD.y), __float_as_int(wrD.z), exact ? 1 : 0);
            sp = stackBase + STRIDE, cur = s.root_ref;
            rays++;
        }
    }
#undef RT_S8_NODE_ACTION
#undef TO
#undef TD
    for (int off = 16; off; off >>= 1) rays += __shfl_xor_sync(FULL, rays, off);
    if (lane == 0) atomicAdd(p.counters, (unsigned long long)rays);
}

// accumulator += the images of a launch, in image order, for the pixels of the launch's tiles (renderer.cpp:124: one float add per
// channel per sample, frame after frame, pass after pass - the order of the reference's Tick sequence).  One CTA = one 16 x 16 tile.
// `accum` may live on another GPU (peer-mapped: tile-sharded multi-GPU renders write their tiles straight into rank 0's image).
__global__ void __launch_bounds__(256) k_sum_frames(float4* __restrict__ accum, const float4* __restrict__ images, const int nImages,
    const int W, const int H, const int tilesX, const int tileBegin, const int tileStep, const int nTiles, const int compact)
{
    // image stride and this thread's pixel inside an image: W x H images, or compact ones holding the job's tiles only
    const size_t imagePixels = compact ? (size_t)nTiles * 256 : (size_t)W * H;
    for (int k = blockIdx.x; k < nTiles; k += gridDim.x)
    {
        const int tile = tileBegin + k * tileStep;
        const int x = (tile % tilesX) * 16 + (threadIdx.x & 15), y = (tile / tilesX) * 16 + (threadIdx.x >> 4);
        const size_t pixel = x + (size_t)y * W;
        float4 a = accum[pixel];
        const float4* f = images + (compact ? (size_t)k * 256 + threadIdx.x : pixel);
        int i = 0;
        for (; i + 4 <= nImages; i += 4)
        {
            // four loads in flight, added in order
            const float4 f0 = __ldcs(f), f1 = __ldcs(f + imagePixels), f2 = __ldcs(f + 2 * imagePixels), f3_ = __ldcs(f + 3 * imagePixels);
            a.x += f0.x, a.y += f0.y, a.z += f0.z;
            a.x += f1.x, a.y += f1.y, a.z += f1.z;
            a.x += f2.x, a.y += f2.y, a.z += f2.z;
            a.x += f3_.x, a.y += f3_.y, a.z += f3_.z;
            f += 4 * imagePixels;
        }
        for (; i < nImages; i++, f += imagePixels)
        {
            const float4 f0 = __ldcs(f);
            a.x += f0.x, a.y += f0.y, a.z += f0.z;
        }
        accum[pixel] = a;
    }
}

// Renderer::ClearAccumulator for the tiles of one shard (rt_renderer_clear of a tile-sharded renderer)
__global__ void __launch_bounds__(256) k_clear_tiles(float4* __restrict__ accum, const int W, const int tilesX, const int tileBegin, const int tileStep, const int nTiles)
{
    for (int k = blockIdx.x; k < nTiles; k += gridDim.x)
    {
        const int tile = tileBegin + k * tileStep;
        const int x = (tile % tilesX) * 16 + (threadIdx.x & 15), y = (tile / tilesX) * 16 + (threadIdx.x >> 4);
        accum[x + (size_t)y * W] = make_float4(0, 0, 0, 0);
    }
}

} // namespace rtb
