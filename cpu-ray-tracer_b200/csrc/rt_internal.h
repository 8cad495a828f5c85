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
    size_t node_count = 0, tri_count = 0, inst_count = 0;
    // renderers created on this scene: rt_scene_destroy while some exist only marks the scene, the last rt_renderer_destroy frees it
    std::atomic<int> renderers{0};
    std::atomic<bool> destroy_requested{false};
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
    int* next_fetch_counter() { return fetch_counters + (fetch_next.fetch_add(1) % FETCH_RING); }
};
