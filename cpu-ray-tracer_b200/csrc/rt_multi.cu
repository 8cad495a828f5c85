// rt_multi.cu — Renderer::Tick on several GPUs of one process (include/rt_b200.h, "rt_multi_renderer").
//
// The reference hands the 16x16 tile jobs of a frame to a thread pool (3. PathTracer/renderer.cpp:144-168); tiles never
// interact (each ProcessTile owns its 256 accumulator entries and its own RNG stream, :117-131).  Here the pool is the GPUs of
// the box: scene replicated per device, device k renders tiles k, k + n, ... of every frame of the call with the ordinary
// single-GPU renderer (tile_begin = k, tile_step = n), and all renderers accumulate into ONE image that lives on devices[0]:
// the other devices reach it through peer-mapped memory, so the only inter-GPU traffic is the final frame-ordered sum of each
// shard's tiles (k_sum_frames: 16 B per pixel per call, written over NVLink).  No collective, no host copy; disjoint pixels
// make the image bit-identical to a one-GPU render.
#include <vector>

#include "rt_internal.h"

using namespace rtb;

struct rt_multi_renderer {
    std::vector<int> devices;
    std::vector<rt_scene*> scenes;
    std::vector<rt_renderer*> renderers;
    int integrator = RT_INTEGRATOR_PATH;
};

extern "C" {

void rt_multi_renderer_destroy(rt_multi_renderer* m)
{
    if (!m) return;
    // shards first (they write into renderers[0]'s accumulator), the owner of the image last
    for (size_t k = m->renderers.size(); k-- > 0;) rt_renderer_destroy(m->renderers[k]);
    for (rt_scene* s : m->scenes) rt_scene_destroy(s);
    delete m;
}

rt_status rt_multi_renderer_create(const rt_scene_desc* desc, uint32_t scene_flags, const int* devices, int n,
                                   const rt_render_params* params, rt_multi_renderer** out)
{
    if (!desc || !devices || !params || !out || n < 1) { set_error("rt_multi_renderer_create: bad argument"); return RT_ERR_INVALID; }
    *out = nullptr;
    if (params->integrator != RT_INTEGRATOR_PATH) { set_error("rt_multi_renderer_create: the path tracer only (a Whitted frame is one short launch sequence on one GPU)"); return RT_ERR_UNSUPPORTED; }
    if (params->tile_begin != 0 || params->tile_end > 0 || params->tile_step > 1) { set_error("rt_multi_renderer_create: the tile range is split by the library"); return RT_ERR_INVALID; }
    const int have = rt_device_count();
    for (int k = 0; k < n; k++)
    {
        if (devices[k] < 0 || devices[k] >= have) { set_error("rt_multi_renderer_create: no such CUDA device (there is no CPU fallback)"); return RT_ERR_NO_DEVICE; }
        for (int j = 0; j < k; j++)
            if (devices[j] == devices[k]) { set_error("rt_multi_renderer_create: a device is listed twice"); return RT_ERR_INVALID; }
    }
    for (int k = 1; k < n; k++)
    {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[k], devices[0]) != cudaSuccess || !can)
        {
            cudaGetLastError();
            set_error("rt_multi_renderer_create: device " + std::to_string(devices[k]) + " cannot map the memory of device " + std::to_string(devices[0]) + " (no peer access)");
            return RT_ERR_UNSUPPORTED;
        }
    }
    rt_multi_renderer* m = new rt_multi_renderer();
    m->integrator = params->integrator;
    auto fail = [&](rt_status st) { rt_multi_renderer_destroy(m); return st; };
    const int tiles = (params->width / 16) * (params->height / 16);
    for (int k = 0; k < n; k++)
    {
        rt_scene* s = nullptr;
        rt_status st = rt_scene_create(desc, devices[k], scene_flags, &s);
        if (st != RT_OK) return fail(st);
        m->devices.push_back(devices[k]), m->scenes.push_back(s);
        rt_render_params p = *params;
        p.tile_begin = k, p.tile_end = tiles, p.tile_step = n;
        rt_renderer* r = nullptr;
        if ((st = rt_renderer_create(s, &p, &r)) != RT_OK) return fail(st);
        m->renderers.push_back(r);
        if (k > 0)
        {
            if (cudaSetDevice(devices[k]) != cudaSuccess) { set_error("cudaSetDevice failed"); return fail(RT_ERR_CUDA); }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cuda_ok(e, "cudaDeviceEnablePeerAccess"); return fail(RT_ERR_CUDA); }
            cudaGetLastError();
            if ((st = rt_renderer_set_accumulator(r, rt_renderer_device_accumulator(m->renderers[0]))) != RT_OK) return fail(st);
        }
    }
    *out = m;
    return RT_OK;
}

int rt_multi_renderer_device_count(const rt_multi_renderer* m) { return m ? (int)m->devices.size() : 0; }

#define RT_MULTI_EACH(CALL)                                            \
    if (!m) return RT_ERR_INVALID;                                     \
    for (rt_renderer* r : m->renderers)                                \
    {                                                                  \
        const rt_status st = (CALL);                                   \
        if (st != RT_OK) return st;                                    \
    }                                                                  \
    return RT_OK;

rt_status rt_multi_renderer_set_camera(rt_multi_renderer* m, const rt_camera* cam) { RT_MULTI_EACH(rt_renderer_set_camera(r, cam)) }
rt_status rt_multi_renderer_set_passes(rt_multi_renderer* m, int passes) { RT_MULTI_EACH(rt_renderer_set_passes(r, passes)) }
rt_status rt_multi_renderer_clear(rt_multi_renderer* m) { RT_MULTI_EACH(rt_renderer_clear(r)) }
rt_status rt_multi_renderer_render(rt_multi_renderer* m, int first_spp, int count, int stride) { RT_MULTI_EACH(rt_renderer_render(r, first_spp, count, stride)) }
rt_status rt_multi_renderer_sync(rt_multi_renderer* m) { RT_MULTI_EACH(rt_renderer_sync(r)) }
rt_status rt_multi_renderer_reset_counters(rt_multi_renderer* m) { RT_MULTI_EACH(rt_renderer_reset_counters(r)) }

rt_status rt_multi_renderer_read_accumulator(rt_multi_renderer* m, float* host_rgba)
{
    const rt_status st = rt_multi_renderer_sync(m);
    if (st != RT_OK) return st;
    return rt_renderer_read_accumulator(m->renderers[0], host_rgba);
}

rt_status rt_multi_renderer_read_pixels(rt_multi_renderer* m, float scale, uint32_t* host_rgb8)
{
    const rt_status st = rt_multi_renderer_sync(m);
    if (st != RT_OK) return st;
    return rt_renderer_read_pixels(m->renderers[0], scale, host_rgb8);
}

rt_status rt_multi_renderer_get_counters(rt_multi_renderer* m, rt_counters* out)
{
    if (!m || !out) return RT_ERR_INVALID;
    rt_counters sum = {};
    for (rt_renderer* r : m->renderers)
    {
        rt_counters c;
        const rt_status st = rt_renderer_get_counters(r, &c);
        if (st != RT_OK) return st;
        sum.extension_rays += c.extension_rays, sum.shadow_rays += c.shadow_rays, sum.paths += c.paths;
        sum.wavefront_iterations += c.wavefront_iterations, sum.kernel_launches += c.kernel_launches;
    }
    *out = sum;
    return RT_OK;
}

} // extern "C"
