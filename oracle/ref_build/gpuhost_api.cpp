// C driver over the C++ drop-in adapters (TEST INFRASTRUCTURE ONLY; appended to the unity translation
// unit by build_ref.py for the libgpuhost_* variants, after ref_api.cpp).
//
// It is what a maintainer's main() would do after switching to the GPU core: construct
// rtb200::GpuRenderer<SceneT, INTEGRATOR> instead of Renderer — the scene XML, OBJ and texture loading
// and the SAH / TLAS builders that run inside are the reference's own, unchanged — and call
// Init / Tick.  tests/test_cpp_adapters.py compares the result with the reference's Renderer (exposed by
// ref_api.cpp in the same library) on the same scene.
#include "rt_b200_adapters.h"

#if defined(REF_INTEGRATOR_PT)
typedef rtb200::GpuRenderer<REF_SCENE_TYPE, RT_INTEGRATOR_PATH> GhRenderer;
#else
typedef rtb200::GpuRenderer<REF_SCENE_TYPE, RT_INTEGRATOR_WHITTED> GhRenderer;
#endif

static GhRenderer* g_gh = nullptr;
static std::string g_gh_error;

extern "C" {

const char* gh_last_error() { return g_gh_error.c_str(); }

// same contract as ref_create: workdir's parent holds `assets/`
int gh_create( const char* scene_xml, const char* workdir, int width, int height, int device )
{
	try
	{
		if (workdir && workdir[0] && chdir( workdir ) != 0) return -2;
		g_ref_scrwidth = width, g_ref_scrheight = height;
		g_gh = new GhRenderer( scene_xml, device );
		g_gh->screen = new Surface( width, height );
		g_gh->Init();
		return 0;
	}
	catch (const std::exception& e)
	{
		g_gh_error = e.what();
		return -1;
	}
}

// the same on several GPUs of this process (path tracer: rt_multi_renderer behind GpuRenderer::Tick)
int gh_create_multi( const char* scene_xml, const char* workdir, int width, int height, const int* devices, int n )
{
	try
	{
		if (workdir && workdir[0] && chdir( workdir ) != 0) return -2;
		g_ref_scrwidth = width, g_ref_scrheight = height;
		g_gh = new GhRenderer( scene_xml, std::vector<int>( devices, devices + n ) );
		g_gh->screen = new Surface( width, height );
		g_gh->Init();
		return 0;
	}
	catch (const std::exception& e)
	{
		g_gh_error = e.what();
		return -1;
	}
}

void gh_destroy() { delete g_gh; g_gh = nullptr; }

// out: instances, meshes resident on the device, fat nodes, triangle slots, geometry bytes, devices rendering
int gh_scene_info( unsigned long long* out )
{
	try
	{
		const rt_scene_info i = g_gh->scene.Info();
		out[0] = i.instances, out[1] = i.meshes, out[2] = i.fat_nodes, out[3] = i.triangle_slots, out[4] = i.bytes_geometry;
		out[5] = (unsigned long long)g_gh->DeviceCount() * (g_gh->MultiHandle() ? 1 : 0);
		return 0;
	}
	catch (const std::exception& e) { g_gh_error = e.what(); return -1; }
}

void gh_set_camera( const float* pos, const float* target )
{
	g_gh->camera.SetCameraState( float3( pos[0], pos[1], pos[2] ), float3( target[0], target[1], target[2] ) );
}

int gh_tick( int frames )
{
	try { for (int i = 0; i < frames; i++) g_gh->Tick( 0 ); return 0; }
	catch (const std::exception& e) { g_gh_error = e.what(); return -1; }
}

int gh_render( int frames )
{
	try { g_gh->Render( frames ); return 0; }
	catch (const std::exception& e) { g_gh_error = e.what(); return -1; }
}

const float* gh_accumulator() { return (const float*)g_gh->accumulator; }
const unsigned int* gh_screen() { return g_gh->screen->pixels; }
int gh_spp() { return g_gh->spp; }
void gh_set_depth_limit( int d ) { g_gh->depthLimit = d; }
void gh_set_passes( int p ) { g_gh->passes = p; }
int gh_triangle_count() { return g_gh->scene.GetTriangleCount(); }

int gh_reset( int spp )
{
	try { g_gh->ClearAccumulator(); g_gh->spp = spp; return 0; }
	catch (const std::exception& e) { g_gh_error = e.what(); return -1; }
}

// BaseScene::FindNearest through the adapter: batched, and (single != 0) one virtual call per ray
int gh_find_nearest( int n, const float* O, const float* D, const float* tmax, int single,
	float* t, float* u, float* v, int* objIdx, int* triIdx )
{
	try
	{
		std::vector<Ray> rays( n );
		for (int i = 0; i < n; i++) api_ray( rays[i], O + 3 * i, D + 3 * i, tmax[i] );
		if (single) { Tmpl8::BaseScene* s = &g_gh->scene; for (int i = 0; i < n; i++) s->FindNearest( rays[i] ); }
		else g_gh->scene.FindNearest( rays.data(), (size_t)n );
		for (int i = 0; i < n; i++)
			t[i] = rays[i].t, u[i] = rays[i].barycentric.x, v[i] = rays[i].barycentric.y, objIdx[i] = rays[i].objIdx, triIdx[i] = rays[i].triIdx;
		return 0;
	}
	catch (const std::exception& e) { g_gh_error = e.what(); return -1; }
}

int gh_is_occluded( int n, const float* O, const float* D, const float* tmax, unsigned char* out )
{
	try
	{
		std::vector<Ray> rays( n );
		for (int i = 0; i < n; i++) api_ray( rays[i], O + 3 * i, D + 3 * i, tmax[i] );
		g_gh->scene.IsOccluded( rays.data(), (size_t)n, out );
		return 0;
	}
	catch (const std::exception& e) { g_gh_error = e.what(); return -1; }
}

} // extern "C"
