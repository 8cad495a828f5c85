// rt_internal.h — host-side objects behind the opaque handles of include/rt_b200.h
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "rt_device.cuh"

namespace rtb {

void set_error(const std::string& msg);
bool cuda_ok(cudaError_t e, const char* what);

#define RT_CUDA(call)                                            \
    do {                                                         \
        if (!rtb::cuda_ok((call), #call)) return RT_ERR_CUDA;    \
    } while (0)

// Result of the device SAH build (rt_build.cu): the reference-layout arrays, left in device memory
struct DeviceBvh {
    rt_tri* tris = nullptr;        // the uploaded 112-byte triangles
    rt_bvh_node* nodes = nullptr;  // `total` nodes, reference numbering
    uint32_t* idx = nullptr;       // triangleIndices
    uint32_t n = 0, total = 0;
    int depth = 0;                 // levels of the tree (a single leaf = 1)
    double ms = 0;
    void release();
};
rt_status build_bvh_on_device(int device, const rt_tri* host_tris, uint32_t n, DeviceBvh& out);

// One mesh of a BVH-kind scene on the device (several BLAS instances may share it): where its fat nodes, triangle records and
// shading records live.  rt_scene_refit works on these ranges.
struct Geometry {
    int fatBase = 0, fatCount = 0;   // interior nodes: nodes[4 * fatBase .. 4 * (fatBase + fatCount))
    int triBase = 0;                 // first triangle slot = first shading record
    uint32_t triCount = 0;
    int rootRef = 0;
};

// rt_construct.cu: scene construction steps that run on the device
rt_status layout_built_bvh(const DeviceBvh& b, int fatBase, int triBase, float4* nodes, float4* tris, float4* shade, cudaStream_t stream);
rt_status build_tlas_on_device(int device, const float* d_world_bounds, uint32_t n, rt_tlas_node32* d_out, int* d_depth, cudaStream_t stream);
rt_status layout_built_tlas(const rt_tlas_node32* d_tlas, uint32_t n, int fatBase, float4* nodes, cudaStream_t stream);
void world_bounds_of(const float* root_min, const float* root_max, const float* T, float* out6);

} // namespace rtb

// Runs CALL with `A` = the accelerator id the scene kind selects (kernel template argument).
#define RT_FOR_ACCEL(kind, CALL)                                                           \
    switch (kind)                                                                          \
    {                                                                                      \
    case RT_SCENE_FLAT_KDTREE: { constexpr int A = rtb::ACCEL_KD; CALL; } break;               \
    case RT_SCENE_FLAT_GRID: { constexpr int A = rtb::ACCEL_GRID; CALL; } break;               \
    case RT_SCENE_TLAS_KDTREE: { constexpr int A = rtb::ACCEL_TLAS_KD; CALL; } break;          \
    case RT_SCENE_TLAS_GRID: { constexpr int A = rtb::ACCEL_TLAS_GRID; CALL; } break;          \
    default: { constexpr int A = rtb::ACCEL_BVH; CALL; } break;                                \
    }

struct rt_scene {
    int device = 0;
    uint32_t flags = 0;
    rtb::DScene d = {};
    // owned device allocations
    float4* nodes = nullptr;
    float4* tris = nullptr;
    float4* inst = nullptr;
    float4* shade = nullptr;
    float4* inst_shade = nullptr;
    float4* kd_nodes = nullptr;   // RT_SCENE_FLAT_KDTREE
    int2* grid_cells = nullptr;   // RT_SCENE_FLAT_GRID
    float4* grid_params = nullptr;
    int* obj_material = nullptr;
    rtb::DMaterial* materials = nullptr;
    rtb::DTexture* textures = nullptr;
    uint32_t* tex_pixels = nullptr;
    size_t node_count = 0, tri_count = 0, inst_count = 0, mesh_count = 0;
    // renderers created on this scene: rt_scene_destroy while some exist only marks the scene, the last rt_renderer_destroy frees it
    std::atomic<int> renderers{0};
    std::atomic<bool> destroy_requested{false};
    // BVH kinds: the meshes and which one every BLAS instance uses (rt_scene_refit); TLAS: where its fat nodes live
    std::vector<rtb::Geometry> geometries;
    std::vector<int> blas_geometry;
    std::vector<float> blas_T;          // 16 floats per instance (world bounds after a refit)
    std::vector<float> root_boxes;      // 6 floats per mesh: the box of its BVH root as the builder left it (single-leaf meshes have no fat node)
    int tlas_fat_base = 0, tlas_fat_count = 0;
    int max_blas_depth = 0;
    int stack_entries = 0; // most far children one ray can have pending (bound from the tree depths, rt_scene_create)
    size_t bytes_geometry = 0, bytes_textures = 0;
    // scratch for the host-buffer entry points (rt_find_nearest / rt_is_occluded)
    std::mutex scratch_mutex;
    void* scratch_in = nullptr;
    void* scratch_out = nullptr;
    size_t scratch_in_bytes = 0, scratch_out_bytes = 0;
    cudaStream_t stream = nullptr;
    // queue-fetch counters of the persistent traversal kernels: a ring, so that launches in flight on
    // different streams do not share one
    static constexpr int FETCH_RING = 256;
    int* fetch_counters = nullptr;
    std::atomic<unsigned> fetch_next{0};
    bool persistent = true, voted = false;
    // north_star (b) "L2-resident top levels": bytes of the L2 set aside for the fat-node array (access-policy window on the
    // streams that traverse; RT_B200_L2_PERSIST_MB, 0 = off).  Node fetches then keep their lines against the streaming ray /
    // hit / frame-image traffic; for node arrays larger than the set-aside the window persists that fraction of the lines.
    size_t l2_persist_bytes = 0;
    void apply_l2_policy(cudaStream_t stream) const;
    int* next_fetch_counter() { return fetch_counters + (fetch_next.fetch_add(1) % FETCH_RING); }
};
