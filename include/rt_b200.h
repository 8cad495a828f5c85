/* rt_b200.h — C-ABI of the B200-native ray core (drop-in for the hot path of willake/cpu-ray-tracer).
 *
 * The reference has no plugin/FFI boundary; the seam is two C++ class surfaces:
 *   - BaseScene virtuals  FindNearest / IsOccluded / GetHitInfo / GetSkyColor / GetLightPos / GetLightColor
 *     (infra/scene/base_scene.h:16-32, implemented by FileScene file_scene.cpp:142-214 and TLASFileScene
 *     tlas_file_scene.cpp:173-260), and
 *   - Renderer : TheApp    Init / Tick / Trace / Sample + public `accumulator`, `camera`, `spp`, `passes`,
 *     `depthLimit` (2. WhittedStyle/renderer.h:41-61, 3. PathTracer/renderer.h:29-53).
 * This header is what a binding of those two surfaces calls.  Everything is plain C: pointers, sizes,
 * POD structs, int status codes; no exceptions cross it.  Host-side scene/OBJ/XML/image loading and
 * the SAH/TLAS builders stay where they are in the reference; after construction the host hands the
 * builders' own arrays over unchanged (the POD structs below have the reference's exact memory layout,
 * so `acc.bvhNodes.data()` etc. can be passed without conversion).
 *
 * All rt_* entry points are implemented in CUDA for sm_100a (cpu-ray-tracer_b200/csrc); there is no CPU
 * fallback: without a CUDA device every compute call returns RT_ERR_NO_DEVICE.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_B200_ABI_VERSION 5

typedef int rt_status;
enum {
    RT_OK = 0,
    RT_ERR_INVALID = -1,    /* null / inconsistent argument */
    RT_ERR_NO_DEVICE = -2,  /* no CUDA device: the product path never falls back to the CPU */
    RT_ERR_CUDA = -3,       /* a CUDA call failed; rt_last_error() has the text */
    RT_ERR_UNSUPPORTED = -4
};

/* ---- flattened scene: the reference builders' own arrays ------------------------------------ */

/* = Tmpl8::BVHNode, infra/blas_bvh.h:13-20 (32 bytes). tri_count > 0 => leaf, left_first = first
 * entry in tri_indices; else left_first = index of the left child, right child = left_first + 1. */
typedef struct rt_bvh_node {
    float aabb_min[3];
    float aabb_max[3];
    uint32_t left_first;
    uint32_t tri_count;
} rt_bvh_node;

/* = Tri, infra/helper.h:6-26 (112 bytes). */
typedef struct rt_tri {
    float v0[3], v1[3], v2[3];
    float n0[3], n1[3], n2[3];
    float uv0[2], uv1[2], uv2[2];
    float centroid[3];
    int32_t obj_idx;
} rt_tri;

/* = Tmpl8::TLASBVHNode, infra/tlas_bvh.h:7-14 (32 bytes). left_right == 0 => leaf holding `blas`;
 * else children are tlas_nodes[left_right & 0xffff] and tlas_nodes[left_right >> 16]. Node 0 = root. */
typedef struct rt_tlas_node {
    float aabb_min[3];
    uint32_t left_right;
    float aabb_max[3];
    uint32_t blas;
} rt_tlas_node;

/* The same node with 32-bit child indices (ABI v5): the reference packs both children into one 32-bit word, which caps a
 * scene at 32 767 instances (tlas_bvh.h:10; its builder also keeps `int nodeIdx[256]`, tlas_bvh.cpp:21 - README "known
 * issues").  left == 0 => leaf holding BLAS `right`; else children are nodes[left] and nodes[right].  Node 0 = root.
 * Produced by rt_build_tlas; accepted by rt_scene_create through rt_scene_desc.tlas_nodes32. */
typedef struct rt_tlas_node32 {
    float aabb_min[3];
    uint32_t left;
    float aabb_max[3];
    uint32_t right;
} rt_tlas_node32;

/* One bottom-level BVH: BVH (infra/bvh.h:38-41, FileScene::acc) or BLASBVH (infra/blas_bvh.h:44-53).
 * T / inv_T are row-major mat4::cell arrays (identity for the flat FileScene BVH).
 * obj_idx: the BLAS' objIdx written into hits (blas_bvh.cpp:299); -1 for the flat BVH, where the hit
 * takes tri.obj_idx instead (bvh.cpp:220). */
typedef struct rt_blas_desc {
    /* nodes == NULL && tri_indices == NULL (ABI v5): the BVH is built ON THE DEVICE from `tris` (rt_build_bvh's builder:
     * the reference's binned SAH, bit-identical arrays) and laid out for traversal there, without a round trip through
     * host memory; node_count is ignored.  Descriptors with the same `tris` pointer and tri_count share one build. */
    const rt_bvh_node* nodes;
    uint32_t node_count;            /* nodesUsed */
    const rt_tri* tris;
    const uint32_t* tri_indices;
    uint32_t tri_count;
    float T[16];
    float inv_T[16];
    int32_t obj_idx;
    int32_t mat_idx;
} rt_blas_desc;

/* = Tmpl8::Material (template/material.h:36-45); texture = index into rt_scene_desc.textures or -1. */
typedef struct rt_material {
    float reflectivity;
    float refractivity;
    float absorption[3];
    float albedo[3];
    int32_t is_light;
    int32_t texture;
} rt_material;

/* = Tmpl8::Texture (template/texture.h:98-103): packed 0x00RRGGBB texels, row 0 first. */
typedef struct rt_texture {
    const uint32_t* pixels;
    int32_t width, height;
} rt_texture;

/* One node of the reference's KD-tree (KDTreeNode, infra/blas_kdtree.h:16-25; built by KDTree::Subdivide,
 * infra/kdtree.cpp:45-112).  The reference links nodes by pointer and keeps a std::vector<uint> per leaf;
 * the host flattens them once after KDTree::Build (loop shown in INTEGRATION.md): node 0 = root,
 * left / right = node indices (-1 in a leaf), a leaf's triangle list = kd_tri_indices[tri_start .. +tri_count)
 * in the leaf's own order (a triangle straddling a split plane is listed in several leaves). */
typedef struct rt_kd_node {
    float aabb_min[3];
    int32_t left;
    float aabb_max[3];
    int32_t right;
    int32_t split_axis;
    float split_distance;   /* split plane = aabb_min[split_axis] + split_distance (kdtree.cpp:171) */
    uint32_t tri_start, tri_count;
} rt_kd_node;

/* The reference's uniform grid (Grid, infra/grid.h:25-34; built by Grid::Build, infra/grid.cpp:4-60), flattened:
 * cell (x, y, z) has index x + y*res.x + z*res.x*res.y (grid.cpp:52) and lists
 * tri_indices[cell_start[i] .. cell_start[i+1]) in the order Grid::Build pushed them. */
typedef struct rt_grid_desc {
    int32_t resolution[3];
    float cell_size[3];
    float bounds_min[3], bounds_max[3];   /* Grid::localBounds */
    const uint32_t* cell_start;           /* resolution.x*y*z + 1 entries */
    const uint32_t* tri_indices;
    uint32_t index_count;
} rt_grid_desc;

/* Per-BLAS KD-tree / grid of TLASFileScene built with TLAS_USE_KDTree / TLAS_USE_Grid (tlas_file_scene.h:12-14):
 * BLASKDTree (infra/blas_kdtree.h:27-58) or BLASGrid (infra/blas_grid.h:13-41), flattened like the flat-scene forms
 * above; node / cell indices and tri_start are relative to this BLAS' own arrays.  The TLAS above them is the same
 * agglomerative BVH as for TLAS_USE_BVH (tlas_kdtree.cpp:17-70 = tlas_grid.cpp:17-70 = tlas_bvh.cpp:17-70). */
typedef struct rt_blas_accel {
    const rt_kd_node* kd_nodes;     /* RT_SCENE_TLAS_KDTREE */
    uint32_t kd_node_count;
    const uint32_t* kd_tri_indices;
    uint32_t kd_tri_index_count;
    const rt_grid_desc* grid;       /* RT_SCENE_TLAS_GRID */
} rt_blas_accel;

enum {
    RT_SCENE_FLAT = 0,          /* FileScene, USE_BVH */
    RT_SCENE_TLAS = 1,          /* TLASFileScene, TLAS_USE_BVH */
    RT_SCENE_FLAT_KDTREE = 2,   /* FileScene, USE_KDTree (the configuration the reference ships, file_scene.h:10-12) */
    RT_SCENE_FLAT_GRID = 3,     /* FileScene, USE_Grid */
    RT_SCENE_TLAS_KDTREE = 4,   /* TLASFileScene, TLAS_USE_KDTree */
    RT_SCENE_TLAS_GRID = 5      /* TLASFileScene, TLAS_USE_Grid */
};

typedef struct rt_scene_desc {
    int32_t kind;
    const rt_blas_desc* blas;       /* RT_SCENE_FLAT: exactly one */
    uint32_t blas_count;
    const rt_tlas_node* tlas_nodes; /* RT_SCENE_TLAS only */
    uint32_t tlas_node_count;
    /* material index of object (objIdx - 2): models[i]->matIdx (file_scene.cpp:207) or
     * blas[i]->matIdx (tlas_file_scene.cpp:237-240) */
    const int32_t* obj_material;
    uint32_t obj_count;
    const rt_material* materials;
    uint32_t material_count;
    const rt_texture* textures;
    uint32_t texture_count;
    int32_t skydome_texture;        /* FileScene::skydome */
    int32_t floor_texture;          /* primitiveMaterials[1].textureDiffuse */
    float floor_n[3];               /* Plane floor: N, d, invto (file_scene.cpp:16, primitives.h:104-106) */
    float floor_d;
    float floor_invto;
    float light_T[16];              /* Quad light (file_scene.cpp:15-19, primitives.h:324-329) */
    float light_inv_T[16];
    float light_size;               /* half edge */
    float light_color[3];           /* GetLightColor(): (24,24,22) */
    float light_pos[3];             /* GetLightPos() */
    /* ABI v3: the other two accelerators of FileScene.  For these kinds blas[0]
     * carries only tris / tri_count / obj_idx = -1 (KDTree::triangles, Grid::triangles); nodes and tri_indices
     * are ignored. */
    const rt_kd_node* kd_nodes;     /* RT_SCENE_FLAT_KDTREE */
    uint32_t kd_node_count;
    const uint32_t* kd_tri_indices;
    uint32_t kd_tri_index_count;
    const rt_grid_desc* grid;       /* RT_SCENE_FLAT_GRID */
    /* RT_SCENE_TLAS_KDTREE / RT_SCENE_TLAS_GRID: blas_count entries, parallel to blas[] (which then carries tris /
     * tri_count / T / inv_T / obj_idx / mat_idx; nodes and tri_indices are ignored) */
    const rt_blas_accel* blas_accel;
    /* ABI v5, RT_SCENE_TLAS: a TLAS with 32-bit child indices (more than 32 767 instances); used instead of tlas_nodes
     * when not NULL.  With tlas_nodes == NULL and tlas_nodes32 == NULL the TLAS is built ON THE DEVICE: world bounds of
     * every instance as BLASBVH::SetTransform computes them (blas_bvh.cpp:369-373), then the reference's agglomerative
     * clustering (TLASBVH::Build, tlas_bvh.cpp:17-70) - the same tree, node for node, without the 256 / 32 767 caps. */
    const rt_tlas_node32* tlas_nodes32;
    uint32_t tlas_node32_count;
} rt_scene_desc;

/* ---- rays and hits (Ray, template/ray.h:6-41, as plain records) ------------------------------ */

typedef struct rt_ray {
    float O[3];
    float tmax;     /* Ray::t on entry (1e34f = unbounded) */
    float D[3];
    int32_t inside; /* Ray::inside */
} rt_ray;

typedef struct rt_hit {
    float t;            /* Ray::t after FindNearest (tmax if nothing was hit) */
    float u, v;         /* Ray::barycentric */
    int32_t obj_idx;    /* -1 = miss, 0 = light quad, 1 = floor plane, >= 2 = model */
    int32_t tri_idx;    /* Ray::triIdx (-1 unless a triangle was hit) */
    int32_t traversed;  /* Ray::traversed / tested work counters (ray.h:38-39), only when     */
    int32_t tested;     /* the scene was created with RT_SCENE_FLAG_COUNTERS, else 0          */
    int32_t reserved;
} rt_hit;

typedef struct rt_scene rt_scene;       /* immutable after creation; shareable between renderers */
typedef struct rt_renderer rt_renderer; /* owns accumulator + wavefront queues; one stream per renderer */

enum { RT_SCENE_FLAG_COUNTERS = 1 };

const char* rt_last_error(void);
int rt_abi_version(void);
int rt_device_count(void);

/* Upload + re-layout for the GPU (fat BVH2 nodes, leaf-ordered float4 triangles, texel arrays).
 * Replaces: the state FileScene/TLASFileScene hold after their constructors ran
 * (file_scene.cpp:4-62, tlas_file_scene.cpp:4-93). */
rt_status rt_scene_create(const rt_scene_desc* desc, int device, uint32_t flags, rt_scene** out);
void rt_scene_destroy(rt_scene* scene);

/* What the scene occupies on the device (BaseScene::GetTriangleCount is the reference's only such query, base_scene.h:31).
 * meshes < instances means BLAS instances share geometry (true instancing, README "known issues").  ABI v5. */
typedef struct rt_scene_info {
    uint64_t fat_nodes;        /* 64-byte interior-node records (BLAS + TLAS) */
    uint64_t triangle_slots;   /* 48-byte leaf-ordered triangle records */
    uint64_t instances;        /* BLAS instances (1 for a flat scene) */
    uint64_t meshes;           /* distinct meshes resident on the device */
    uint64_t bytes_geometry, bytes_textures;
    int32_t stack_entries;     /* most pending far children one ray can have (<= 64) */
    int32_t max_blas_depth;
} rt_scene_info;
rt_status rt_scene_get_info(const rt_scene* scene, rt_scene_info* out);

/* Structural check of the traversal data AS IT LIES IN DEVICE MEMORY (read back): every child reference of every fat node is
 * in range and reached exactly once from its root (a tree), every leaf's triangle run ends inside the triangle array, every
 * triangle tag indexes a shading record of its own mesh, every instance points at a valid root, and the deepest path keeps
 * the traversal stack within its 64 entries.  With these invariants no traversal kernel can read outside the scene's
 * arrays or overflow its stack (the reference's `BVHNode* stack[64]`, bvh.cpp:227, is unchecked).  Meant for tests and for
 * callers that build or refit on the device; compute-sanitizer is the other tool, where the platform allows it.
 * BVH kinds (RT_SCENE_FLAT / RT_SCENE_TLAS).  RT_ERR_INVALID + rt_last_error() names the first violation.  ABI v5. */
rt_status rt_scene_validate(rt_scene* scene);

/* Batched BaseScene::FindNearest (file_scene.cpp:170-175 / tlas_file_scene.cpp:201-206):
 * light quad, floor plane, then the accelerator: BVH (bvh.cpp:224-288), TLAS (tlas_bvh.cpp:83-111), KD-tree
 * (kdtree.cpp:144-210) or grid (grid.cpp:94-161).  Host buffers; H2D + kernel + D2H inside the call. */
rt_status rt_find_nearest(rt_scene* scene, const rt_ray* rays, rt_hit* hits, size_t n);
/* Batched BaseScene::IsOccluded (file_scene.cpp:177-187 / tlas_file_scene.cpp:208-218). */
rt_status rt_is_occluded(rt_scene* scene, const rt_ray* rays, uint8_t* occluded, size_t n);
/* Same, buffers already resident on the scene's device; `stream` is a cudaStream_t (NULL = default). */
rt_status rt_find_nearest_device(rt_scene* scene, const rt_ray* d_rays, rt_hit* d_hits, size_t n, void* stream);
rt_status rt_is_occluded_device(rt_scene* scene, const rt_ray* d_rays, uint8_t* d_occluded, size_t n, void* stream);
/* rt_find_nearest_device with a hint about the batch (ABI v5).  The reference's callers know what they trace - primary rays from
 * the camera (Renderer::Tick) or bounce rays (Sample / Trace recursion) - and the two want different kernels: RT_RAYS_INCOHERENT runs
 * the traversal that votes one action per warp iteration (+3..13 % on bounce rays, slower on camera rays).  Results are identical. */
enum { RT_RAYS_DEFAULT = 0, RT_RAYS_INCOHERENT = 1 };
rt_status rt_find_nearest_device_ex(rt_scene* scene, const rt_ray* d_rays, rt_hit* d_hits, size_t n, void* stream, uint32_t ray_flags);

/* ---- camera (template/camera.h) -------------------------------------------------------------- */

typedef struct rt_camera {
    float pos[3];
    float top_left[3], top_right[3], bottom_left[3];
} rt_camera;

/* Camera::Camera() (camera.h:12-21) for a width x height screen. */
void rt_camera_default(rt_camera* cam, int width, int height);
/* Camera::SetCameraState (camera.h:61-73). */
void rt_camera_look_at(rt_camera* cam, const float pos[3], const float target[3], int width, int height);

/* ---- renderer (Renderer::Init / Tick / accumulator) ----------------------------------------- */

enum { RT_INTEGRATOR_WHITTED = 0 /* 2. WhittedStyle */, RT_INTEGRATOR_PATH = 1 /* 3. PathTracer */ };
enum {
    /* the reference's RNG: one xorshift32 stream per 16x16 tile per frame, seeded
     * InitSeed(tx + ty*W + spp*1799), carried serially through the tile's 256 pixels
     * (3. PathTracer/renderer.cpp:117-131).  Every (tile, frame) stream is one wavefront slot. */
    RT_SEED_REFERENCE_TILE = 0,
    /* one independent stream per pixel per frame (not the reference's sequence; same estimator): no serial chain through
     * a tile, so every pixel of every frame is its own unit of parallelism */
    RT_SEED_PER_PIXEL = 1
};

typedef struct rt_render_params {
    int32_t integrator;
    int32_t width, height;  /* SCRWIDTH / SCRHEIGHT (camera.h:4-5) */
    int32_t depth_limit;    /* Renderer::depthLimit, reference default 5 */
    float epsilon;          /* EPSILON, renderer.h:12, reference 0.001f */
    int32_t seed_mode;
    int32_t tile_begin;     /* path tracer: render tiles [tile_begin, tile_end) of the          */
    int32_t tile_end;       /* (W/16)x(H/16) row-major tile grid; tile_end <= 0 = all tiles      */
    int32_t max_frames_in_flight; /* wavefront width = tiles x frames in flight; 0 = library default */
    int32_t schedule;       /* path tracer, RT_SEED_REFERENCE_TILE: how the (tile, frame) streams are advanced */
    /* Path tracer, one-Tick-per-call use (rt_renderer_render with count == 1): render this many frames ahead in
     * one launch, each into its own image, and add them to the accumulator one per call, in spp order.  The
     * accumulator after every call is exactly what that many Renderer::Tick calls produce (frames depend only
     * on scene, camera and their spp counter, 3. PathTracer/renderer.cpp:120), but a frame costs ~1 ms instead
     * of the ~25 ms a single 256-pixel RNG chain per tile takes on its own.  A camera change discards the
     * frames not yet shown.  0 / 1 = off.  Counters include the rays of frames rendered ahead. */
    int32_t lookahead_frames;
    /* Renderer::passes (3. PathTracer/renderer.h:50, the UI's "spp" slider 1..4): samples per pixel per Tick, drawn
     * consecutively from the tile's stream (renderer.cpp:123-126); a Tick then advances spp by `passes`, so the frames of
     * a fresh renderer are first_spp = 1, stride = passes.  0 = 1.  ABI v3. */
    int32_t passes;
    /* Path tracer: render every tile_step-th tile of [tile_begin, tile_end) (0 = 1).  N processes with tile_begin = rank,
     * tile_step = N split a frame into interleaved tiles, which balances better than contiguous ranges when cost varies
     * across the image.  ABI v3. */
    int32_t tile_step;
} rt_render_params;

enum {
    RT_SCHEDULE_AUTO = 0,      /* library default (streams) */
    /* global wavefront: generate / extend / shade kernels over compacted SoA queues, every stream
     * advances one ray per launch pair */
    RT_SCHEDULE_WAVEFRONT = 1,
    /* one persistent kernel, every lane runs its stream to completion and then pulls the next one:
     * no global lock-step, path state in registers */
    RT_SCHEDULE_STREAMS = 2
};

void rt_render_params_default(rt_render_params* p, int integrator, int width, int height);

typedef struct rt_counters {
    uint64_t extension_rays;  /* FindNearest queries */
    uint64_t shadow_rays;     /* IsOccluded queries */
    uint64_t paths;           /* primary samples */
    uint64_t wavefront_iterations;
    uint64_t kernel_launches;
} rt_counters;

rt_status rt_renderer_create(rt_scene* scene, const rt_render_params* params, rt_renderer** out);
void rt_renderer_destroy(rt_renderer* r);
/* Run on `stream` (cudaStream_t) instead of the renderer's own stream. */
rt_status rt_renderer_set_stream(rt_renderer* r, void* stream);
/* Accumulate into a caller-owned device buffer of width*height float4 (e.g. a torch tensor that is
 * later reduced with NCCL) instead of the renderer's own. */
rt_status rt_renderer_set_accumulator(rt_renderer* r, void* d_accumulator);
rt_status rt_renderer_set_camera(rt_renderer* r, const rt_camera* cam);
/* Public member Renderer::passes; takes effect with the next rt_renderer_render.  1 <= passes <= 8. */
rt_status rt_renderer_set_passes(rt_renderer* r, int passes);
/* Renderer::ClearAccumulator (3. PathTracer/renderer.cpp:15-18). */
rt_status rt_renderer_clear(rt_renderer* r);
/* `count` calls of Renderer::Tick (one frame each), for the frames whose reference `spp` counter
 * would be first_spp, first_spp + stride, ... (a fresh reference renderer starts at spp = 1 and adds
 * `passes` = 1 per frame, 3. PathTracer/renderer.h:50, renderer.cpp:167).  Asynchronous.
 * Whitted ignores first_spp/stride and overwrites the accumulator, as the reference does. */
rt_status rt_renderer_render(rt_renderer* r, int first_spp, int count, int stride);
rt_status rt_renderer_sync(rt_renderer* r);
/* Device -> host copy of the float4 accumulator (synchronises). */
rt_status rt_renderer_read_accumulator(rt_renderer* r, float* host_rgba);
/* screen->pixels as the reference would display it: accumulator * scale through RGBF32_to_RGB8
 * (template/precomp.h:325-341; scale = 1/(spp+passes) for the path tracer, 1 for Whitted). */
rt_status rt_renderer_read_pixels(rt_renderer* r, float scale, uint32_t* host_rgb8);
void* rt_renderer_device_accumulator(rt_renderer* r);

/* Tile-sharded rendering by several PROCESSES (one per GPU, e.g. torchrun): the process that owns the image exports its
 * accumulator as a CUDA IPC handle (RT_IPC_HANDLE_BYTES opaque bytes, sent to the other ranks by any means); the others
 * import it, after which their renderers accumulate straight into that memory over NVLink (peer-mapped stores of the
 * frame-ordered sum, k_sum_frames) - the image is complete after a barrier, no reduce / gather, and bit-identical to a
 * one-GPU render because interleaved tile shards (tile_begin = rank, tile_step = ranks) touch disjoint pixels.
 * rt_renderer_clear of a tile-sharded renderer clears only its own tiles.  ABI v4. */
#define RT_IPC_HANDLE_BYTES 64
rt_status rt_renderer_export_accumulator(rt_renderer* r, void* handle_out);
rt_status rt_renderer_import_accumulator(rt_renderer* r, const void* handle);
rt_status rt_renderer_get_counters(rt_renderer* r, rt_counters* out);
rt_status rt_renderer_reset_counters(rt_renderer* r);

/* ---- Renderer::Tick on several GPUs of ONE process ---------------------------------------------------------
 * The tile jobs of a frame are independent (3. PathTracer/renderer.cpp:144-168 hands them to a job pool); here the pool is
 * the GPUs of the box: the scene is replicated on `n` devices, device k renders tiles k, k + n, ... of every frame
 * (interleaved shards balance the cost that varies across the image) and its frame-ordered sum writes them straight
 * into the accumulator on devices[0] through peer-mapped memory over NVLink.  No reduce, no gather; the accumulator is
 * bit-identical to a one-GPU render of the same frames.  Path tracer only (a Whitted frame is 0.2 ms on one GPU).
 * Calls are asynchronous on every device like rt_renderer_render; rt_multi_renderer_sync / _read_accumulator wait for all.
 * Needs peer access between devices[0] and the others (NVLink / NVSwitch boxes): RT_ERR_UNSUPPORTED otherwise.  ABI v4. */
typedef struct rt_multi_renderer rt_multi_renderer;
rt_status rt_multi_renderer_create(const rt_scene_desc* desc, uint32_t scene_flags, const int* devices, int n,
                                   const rt_render_params* params, rt_multi_renderer** out);
void rt_multi_renderer_destroy(rt_multi_renderer* m);
int rt_multi_renderer_device_count(const rt_multi_renderer* m);
rt_status rt_multi_renderer_set_camera(rt_multi_renderer* m, const rt_camera* cam);
rt_status rt_multi_renderer_set_passes(rt_multi_renderer* m, int passes);
rt_status rt_multi_renderer_clear(rt_multi_renderer* m);
rt_status rt_multi_renderer_render(rt_multi_renderer* m, int first_spp, int count, int stride);
rt_status rt_multi_renderer_sync(rt_multi_renderer* m);
rt_status rt_multi_renderer_read_accumulator(rt_multi_renderer* m, float* host_rgba);
rt_status rt_multi_renderer_read_pixels(rt_multi_renderer* m, float scale, uint32_t* host_rgb8);
/* counters summed over the devices */
rt_status rt_multi_renderer_get_counters(rt_multi_renderer* m, rt_counters* out);
rt_status rt_multi_renderer_reset_counters(rt_multi_renderer* m);

/* Per-stage device time (the reference's own instrumentation is a Timer around the pixel loop,
 * template/precomp.h:146-157, and per-ray traversed/tested counters).  With profiling enabled every
 * kernel launch is bracketed by CUDA events on the renderer's stream; rt_renderer_get_stage_times
 * synchronises, sums the spans per stage, returns them and resets.  Costs ~2 event records per launch,
 * so throughput numbers are taken with profiling off. */
/* RT_STAGE_ACCUMULATE: the frame-ordered sum of a multi-frame launch's sample images into the accumulator (k_sum_frames / k_add_frame) */
enum { RT_STAGE_GENERATE = 0, RT_STAGE_EXTEND = 1, RT_STAGE_SHADE = 2, RT_STAGE_CONNECT = 3, RT_STAGE_ACCUMULATE = 4, RT_STAGE_COUNT = 5 };
typedef struct rt_stage_times {
    double ms[RT_STAGE_COUNT];
    uint64_t launches[RT_STAGE_COUNT];
} rt_stage_times;
rt_status rt_renderer_set_profiling(rt_renderer* r, int enabled);
rt_status rt_renderer_get_stage_times(rt_renderer* r, rt_stage_times* out);
/* Same spans, one by one, in launch order (stage id + milliseconds); call INSTEAD of
 * rt_renderer_get_stage_times (both reset).  *n receives the number written (<= capacity). */
rt_status rt_renderer_get_launch_spans(rt_renderer* r, int32_t* stage, float* ms, size_t capacity, size_t* n);
/* Path tracer: number of queued rays at every wavefront iteration of the last rt_renderer_render batch. */
rt_status rt_renderer_get_queue_history(rt_renderer* r, int32_t* rays_per_iteration, size_t capacity, size_t* n);

/* ---- scene construction on the device ----------------------------------- */

/* BVH::Build / BLASBVH::Build (bvh.cpp:4-24, 45-178 = blas_bvh.cpp:82-257) on the GPU: binned SAH, 8 bins per
 * axis, leaves of <= 2 triangles, the reference's cost model and its index partition.  Host buffers in, host
 * buffers out, in the reference's layouts: nodes_out needs 2n - 1 entries, tri_indices_out n entries.  The
 * output is bit-identical to what the reference's recursive builder produces (node boxes, node numbering and
 * triangle order).  Triangle centroids are read from tris[i].centroid, as the reference's loaders set them.
 * device_ms (optional) receives the kernel time without the host<->device copies. */
rt_status rt_build_bvh(int device, const rt_tri* tris, uint32_t n, rt_bvh_node* nodes_out, uint32_t* tri_indices_out,
                       uint32_t* nodes_used, double* device_ms);

/* TLASBVH::Build (tlas_bvh.cpp:17-70: agglomerative clustering, FindBestMatch = smallest surface area of the union,
 * first candidate wins ties) on the GPU.  world_bounds: n x 6 floats (min.xyz, max.xyz per instance, BLAS order); out: 2n
 * entries in the reference's node order (node 0 = copy of the root, leaves 1..n, merged nodes in creation order) with
 * 32-bit children.  The tree is the reference's, node for node; n is bounded by memory only.  Host buffers. */
rt_status rt_build_tlas(int device, const float* world_bounds, uint32_t n, rt_tlas_node32* out, uint32_t* nodes_used, double* device_ms);

/* BVH::Refit / BLASBVH::Refit (bvh.cpp:26-43 = blas_bvh.cpp:104-121) for the mesh of BLAS `blas_index` after its triangles
 * moved: `tris` (host, tri_count entries, the mesh's own triangle order) replaces positions, normals and uvs; topology
 * (tree, triangleIndices) is kept, every leaf box is recomputed from its triangles (UpdateNodeBounds) and every interior box
 * from its children, bottom up, on the device.  Instances that share the mesh all see the change.
 * flags: 0 reproduces the reference exactly, including its skipped node: the loop `for (i = nodesUsed - 1; i >= 0; i--)
 * if (i != 1)` never updates node 1, the root's LEFT child (nodesUsed starts at 1 here, bvh.h:41), whose box stays stale.
 * RT_REFIT_ALL_NODES also updates node 1.  RT_REFIT_REBUILD_TLAS (TLAS scenes): afterwards recompute the world bounds of
 * every instance (SetTransform, blas_bvh.cpp:369-373) and rebuild the TLAS on the device (TLASBVH::Build); without it
 * the TLAS keeps its boxes, as in the reference, where nothing calls SetTransform again.
 * Renderers of the scene must be idle (rt_renderer_sync) while this runs.  BVH kinds only (RT_SCENE_FLAT / RT_SCENE_TLAS). */
enum { RT_REFIT_ALL_NODES = 1, RT_REFIT_REBUILD_TLAS = 2 };
rt_status rt_scene_refit(rt_scene* scene, uint32_t blas_index, const rt_tri* tris, uint32_t tri_count, uint32_t flags);

/* Reads the traversal layout of one mesh back as reference-layout arrays (parity tests of the device build / refit):
 * nodes_out 2 * tri_count - 1 entries, tri_indices_out tri_count entries.  Node boxes, numbering and triangle order are
 * those of BVH::Build; the root's own box is not stored on the device and is returned as the union of its children. */
rt_status rt_scene_download_bvh(rt_scene* scene, uint32_t blas_index, rt_bvh_node* nodes_out, uint32_t* tri_indices_out, uint32_t* nodes_used);

/* ---- diagnostics ------------------------------------------------------------------------------ */

/* Measured bandwidth of random 64-byte record gathers (the size and alignment of one device BVH node) over
 * a working set of `working_set_bytes` (rounded down to a power of two): the roofline denominator for
 * scenes whose geometry is L2-resident (SURVEY.md section 8d asks for a measured L2 peak).  bypass_l1 != 0
 * uses ld.global.cg.  No reference equivalent (the reference has no memory-hierarchy instrumentation). */
rt_status rt_measure_gather_bandwidth(int device, size_t working_set_bytes, int bypass_l1, double* gb_per_s);
/* Measured bandwidth of STREAMING reads of an L2-resident buffer (every SM, coalesced 128-bit loads, L1 bypassed, 8 loads in
 * flight per thread): the L2 peak the roofline of the L2-resident scenes is quoted against (bench.py `roofline.peak`).
 * working_set_bytes: 32-64 MB keeps the set inside the 126 MB L2.  ABI v4. */
rt_status rt_measure_l2_stream_bandwidth(int device, size_t working_set_bytes, double* gb_per_s);

/* Evaluates, ON THE DEVICE, the transcendental routines the shading kernels use, over host arrays of n arguments:
 * fn = RT_MATH_EXPF: out[i] = expf(a[i])   (Beer's law, renderer.cpp:76-80; b is ignored and may be NULL)
 * fn = RT_MATH_ACOSF: out[i] = acosf(a[i]), fn = RT_MATH_ATAN2F: out[i] = atan2f(a[i], b[i])   (GetSkyColor,
 * file_scene.cpp:142-154).  The kernels call glibc 2.39's routines restated for the device (csrc/rt_glibc_math.cuh), so
 * out[] must equal the host libm's results bit for bit; the parity tests check exactly that.  Test instrumentation:
 * the reference has no equivalent entry point. */
#define RT_MATH_EXPF 0
#define RT_MATH_ACOSF 1
#define RT_MATH_ATAN2F 2
/* fn = RT_MATH_SKY_TEXEL: the sky lookup's texel choice for the direction with azimuth a[i] (radians) and height b[i] in [-1, 1] over a
 * 4096 x 2048 skydome: out[i] = 1 when the fast lookup (CUDA's atan2f / acosf) accepted its texel and it is the texel the glibc
 * routines give, 0 when the lookup was too close to a texel border and was handed to the glibc routines, -1 when the fast lookup
 * accepted a different texel (the parity tests require that this never happens). */
#define RT_MATH_SKY_TEXEL 3
rt_status rt_eval_shading_math(int device, int fn, const float* a, const float* b, float* out, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
