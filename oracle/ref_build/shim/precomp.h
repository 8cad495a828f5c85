// Headless Linux shim for the reference's precompiled header (TEST INFRASTRUCTURE ONLY).
//
// The reference (willake/cpu-ray-tracer) is an MSVC/Win32 project whose template/precomp.h pulls in
// windows.h, GLFW, OpenGL, ImGui and OpenCL.  None of that is on the hot path.  This file is OUR
// replacement for that header: it declares just enough of the template's environment (types, macros,
// Timer, Job/JobManager, TheApp, Surface, no-op ImGui) for the reference's *unmodified* hot-path
// sources (tmplmath, bvh, blas_bvh, tlas_bvh, model, file_scene, tlas_file_scene, renderer) to compile
// with g++ as one unity translation unit.  It contains no reference code; the pieces it mirrors are
// cited (reference file:line) next to each stand-in.
//
// Only oracle/ref_build/build_ref.py uses this file; nothing in the product includes it.
#pragma once
#include <chrono>
#include <fstream>
#include <vector>
#include <list>
#include <string>
#include <thread>
#include <math.h>
#include <cmath>
#include <algorithm>
#include <assert.h>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <stdexcept>
#include <iostream>
#include <memory>
#include <unordered_map>
#include <immintrin.h>
#include <omp.h>

// basic types (template/precomp.h:23-26)
typedef unsigned char uchar;
typedef unsigned int uint;
typedef unsigned short ushort;

using namespace std; // template/precomp.h:36 leaks std the same way

// aligned allocation (template/precomp.h:85-93); C11 aligned_alloc wants size % alignment == 0
#define ALIGN( x ) __attribute__( ( aligned( x ) ) )
static inline void* shim_aligned_malloc( size_t size, size_t align )
{
	size_t padded = (size + align - 1) / align * align;
	return padded == 0 ? nullptr : aligned_alloc( align, padded );
}
#define MALLOC64( x ) shim_aligned_malloc( (size_t)(x), 64 )
#define FREE64( x ) free( x )
#define _aligned_malloc( size, align ) shim_aligned_malloc( (size_t)(size), (align) )
#define _aligned_free( x ) free( x )
#define CHECK_RESULT __attribute__( ( warn_unused_result ) )

// run-time screen size: the reference fixes SCRWIDTH/SCRHEIGHT at compile time (template/camera.h:4-5);
// the oracle library serves several BASELINE configs, so they are globals here.  The expressions that
// use them ("1.0f / SCRWIDTH", "(float)SCRWIDTH / (float)SCRHEIGHT") give the same IEEE result either way.
extern int g_ref_scrwidth, g_ref_scrheight;
#define SCRWIDTH g_ref_scrwidth
#define SCRHEIGHT g_ref_scrheight
extern std::string g_ref_scene_path;

// no-op Dear ImGui (the renderers' UI() and infra/helper.h:122-136 reference it; never called headless)
namespace ImGui
{
	template <class... A> inline bool Checkbox( A... ) { return false; }
	template <class... A> inline bool SliderInt( A... ) { return false; }
	template <class... A> inline bool SliderFloat( A... ) { return false; }
	template <class... A> inline bool InputFloat( A... ) { return false; }
	template <class... A> inline bool Button( A... ) { return false; }
	template <class... A> inline void Text( A... ) {}
	template <class... A> inline void PushID( A... ) {}
	template <class... A> inline void PopID( A... ) {}
	template <class... A> inline void SameLine( A... ) {}
	template <class... A> inline void SetNextItemWidth( A... ) {}
}

// GLFW key codes used by template/camera.h:31-59 (values irrelevant: keys are never down)
enum
{
	GLFW_KEY_A = 65, GLFW_KEY_D = 68, GLFW_KEY_F = 70, GLFW_KEY_R = 82, GLFW_KEY_S = 83, GLFW_KEY_W = 87,
	GLFW_KEY_RIGHT = 262, GLFW_KEY_LEFT = 263, GLFW_KEY_DOWN = 264, GLFW_KEY_UP = 265
};
inline bool IsKeyDown( const uint ) { return false; }
inline bool WindowHasFocus() { return false; }

// FatalError: template/template.cpp shows a message box and exits; texture.h:45 passes a std::string
// through C varargs, so a template is needed to swallow it.
template <class... A> [[noreturn]] inline void FatalError( const char* fmt, A... )
{
	fprintf( stderr, "ref oracle FatalError: %s\n", fmt );
	throw std::runtime_error( std::string( "FatalError: " ) + fmt );
}

namespace Tmpl8
{
	// display surface: only `pixels` is touched by the renderers (template/surface.h:34-59)
	class Surface
	{
	public:
		Surface() = default;
		Surface( int w, int h ) : width( w ), height( h ) { pixels = (uint*)MALLOC64( (size_t)w * h * sizeof( uint ) ); }
		Surface( const char* ) {}
		uint* pixels = 0;
		int width = 0, height = 0;
	};
}
using namespace Tmpl8;

// tmplmath.h:122-123 defines its own global fminf/fmaxf; keep them from colliding with libm's
#define fminf tmpl8_fminf
#define fmaxf tmpl8_fmaxf
#include "tmplmath.h"
#include "common.h"

// timer (template/precomp.h:146-157)
struct Timer
{
	Timer() { reset(); }
	float elapsed() const
	{
		return (float)std::chrono::duration<double>( std::chrono::high_resolution_clock::now() - start ).count();
	}
	void reset() { start = std::chrono::high_resolution_clock::now(); }
	std::chrono::high_resolution_clock::time_point start;
};

// job system (template/precomp.h:159-205, template/template.cpp:361-510): Win32 threads there, one
// worker per logical core popping jobs LIFO.  Jobs are independent (one 16x16 tile each), so an OpenMP
// dynamic loop over the job list is an equivalent schedule.  No 4096-job cap here.
class Job
{
public:
	virtual void Main() = 0;
};
class JobManager
{
public:
	static JobManager* GetJobManager() { static JobManager jm; return &jm; }
	void AddJob2( Job* j ) { jobs.push_back( j ); }
	void RunJobs( bool = true )
	{
		const int n = (int)jobs.size();
#pragma omp parallel for schedule( dynamic, 1 )
		for (int i = n - 1; i >= 0; i--) jobs[i]->Main();
		jobs.clear();
	}
	unsigned int GetNumThreads() { return (unsigned)omp_get_max_threads(); }
	int MaxConcurrent() { return omp_get_max_threads(); }
	std::vector<Job*> jobs;
};

// colour conversion, scalar branch of template/precomp.h:325-341
inline uint RGBF32_to_RGB8( const float4* v )
{
	uint r = (uint)(255.0f * min( 1.0f, v->x ));
	uint g = (uint)(255.0f * min( 1.0f, v->y ));
	uint b = (uint)(255.0f * min( 1.0f, v->z ));
	return (r << 16) + (g << 8) + b;
}

// application base class (template/precomp.h:344-361)
class TheApp
{
public:
	virtual void Init() {}
	virtual void Tick( float deltaTime ) = 0;
	virtual void UI() { uiUpdated = false; }
	virtual void Shutdown() {}
	virtual void MouseUp( int ) {}
	virtual void MouseDown( int ) {}
	virtual void MouseMove( int, int ) {}
	virtual void MouseWheel( float ) {}
	virtual void KeyUp( int ) {}
	virtual void KeyDown( int ) {}
	static inline JobManager* jm = JobManager::GetJobManager();
	Surface* screen = 0;
	bool uiUpdated;
	uint end_of_base_class = 99999;
};

// vendored single-header loaders, exactly as template/precomp.h:370-380 configures them
#define STB_IMAGE_IMPLEMENTATION
#define STBI_NO_PSD
#define STBI_NO_PIC
#define STBI_NO_PNM
#include "stb_image.h"
#define TINYOBJLOADER_IMPLEMENTATION
#include "tiny_obj_loader.h"

// template/precomp.h:382-385
#include "ray.h"
#include "primitives.h"
#include "camera.h"
#include "renderer.h"
