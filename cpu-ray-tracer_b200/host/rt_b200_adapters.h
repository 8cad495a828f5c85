// rt_b200_adapters.h — C++ host side of the drop-in: the reference's Scene / Renderer surface served by
// the CUDA library behind include/rt_b200.h.
//
// This header is compiled INSIDE the reference's translation unit (after its renderer.h), so it sees
// the reference's own types (Tmpl8::Ray, float3, mat4, BVHNode, Tri, TLASBVHNode, BLASBVH, Material,
// Texture, Camera, TheApp).  It contains no reference code; it only walks the containers the
// reference's loaders and builders filled and calls the C-ABI.
//
//   rtb200::GpuScene<FileScene>       replaces FileScene      (infra/scene/file_scene.h; whichever of USE_BVH /
//                                     USE_KDTree - the shipped default - / USE_Grid file_scene.h:10-12 selects)
//   rtb200::GpuScene<TLASFileScene>   replaces TLASFileScene  (infra/scene/tlas_file_scene.h; TLAS_USE_BVH, TLAS_USE_KDTree
//                                     or TLAS_USE_Grid, whichever tlas_file_scene.h:12-14 selects)
//       same BaseScene virtuals (infra/scene/base_scene.h:16-32); FindNearest / IsOccluded run on the
//       GPU (single-ray calls are an n = 1 batch, plus batched overloads); the remaining queries
//       (GetHitInfo, GetAlbedo, GetSkyColor, light) are answered by the host scene it owns, whose
//       loaders / SAH / TLAS builders ran unchanged.
//   rtb200::GpuRenderer<SceneT, INTEGRATOR>   replaces Renderer : TheApp
//       (2. WhittedStyle/renderer.h:41-61, 3. PathTracer/renderer.h:29-53): Init / Tick /
//       ClearAccumulator and the public members accumulator, camera, scene, spp, passes, depthLimit,
//       energy keep their meaning; Tick renders the frame on the GPU and copies the float4 accumulator
//       and screen->pixels back, so template.cpp's frame loop (template.cpp:305-338) works unchanged.
//
// Errors: the reference reports load failures with std::runtime_error (blas_bvh.cpp:11-14); so do the
// adapters for every non-RT_OK status (message = rt_last_error()).  There is no CPU fallback: if the
// library reports RT_ERR_NO_DEVICE the constructor throws.
//
// Requirements on the reference side: read access to TLASBVH::tlasNode / nodesUsed (private at
// tlas_bvh.h:27 — add an accessor or a friend), to Texture::pixels and, for USE_Grid, to Grid::resolution /
// cellSize / gridCells (private at grid.h:25-30).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "rt_b200.h"

namespace rtb200 {

inline void check( rt_status st, const char* what )
{
	if (st != RT_OK) throw std::runtime_error( std::string( what ) + ": " + rt_last_error() );
}

static_assert(sizeof( Tmpl8::BVHNode ) == sizeof( rt_bvh_node ), "BVHNode layout (blas_bvh.h:13-20)");
static_assert(sizeof( Tri ) == sizeof( rt_tri ), "Tri layout (helper.h:6-26)");
static_assert(sizeof( Tmpl8::TLASBVHNode ) == sizeof( rt_tlas_node ), "TLASBVHNode layout (tlas_bvh.h:7-14)");

// The tables rt_scene_desc points to; alive until rt_scene_create has uploaded them.
struct FlattenedScene
{
	std::vector<rt_blas_desc> blas;
	std::vector<int32_t> objMaterial;
	std::vector<rt_material> materials;
	std::vector<rt_texture> textures;
	std::vector<rt_kd_node> kdNodes;      // USE_KDTree / TLAS_USE_KDTree: KDTreeNode graph(s), flattened
	std::vector<uint32_t> altTriIdx;      // KD leaf lists / grid cell lists, concatenated
	std::vector<uint32_t> gridCellStart;  // USE_Grid / TLAS_USE_Grid
	rt_grid_desc grid = {};
	std::vector<rt_grid_desc> blasGrids;  // TLAS_USE_Grid: one per BLAS
	std::vector<rt_blas_accel> blasAccel; // TLAS_USE_KDTree / TLAS_USE_Grid: parallel to blas
	rt_scene_desc desc = {};
};

inline int AddTexture( FlattenedScene& f, const Tmpl8::Texture* tex )
{
	if (!tex || tex->width == 0 || tex->pixels.empty()) return -1;
	rt_texture t;
	t.pixels = (const uint32_t*)tex->pixels.data(), t.width = tex->width, t.height = tex->height;
	f.textures.push_back( t );
	return (int)f.textures.size() - 1;
}

// the parts FileScene and TLASFileScene share: skydome, floor plane, light quad, materials
template <class HostScene> inline void FlattenCommon( HostScene& scene, FlattenedScene& f )
{
	rt_scene_desc& d = f.desc;
	d.skydome_texture = AddTexture( f, &scene.skydome );
	d.floor_texture = AddTexture( f, scene.primitiveMaterials[1].textureDiffuse.get() );
	for (Tmpl8::Material* m : scene.materials)
	{
		rt_material fm = {};
		fm.reflectivity = m->reflectivity, fm.refractivity = m->refractivity;
		fm.absorption[0] = m->absorption.x, fm.absorption[1] = m->absorption.y, fm.absorption[2] = m->absorption.z;
		fm.albedo[0] = m->albedo.x, fm.albedo[1] = m->albedo.y, fm.albedo[2] = m->albedo.z;
		fm.is_light = m->isLight ? 1 : 0;
		fm.texture = AddTexture( f, m->textureDiffuse.get() );
		f.materials.push_back( fm );
	}
	d.floor_n[0] = scene.floor.N.x, d.floor_n[1] = scene.floor.N.y, d.floor_n[2] = scene.floor.N.z;
	d.floor_d = scene.floor.d, d.floor_invto = scene.floor.invto;
	memcpy( d.light_T, scene.light.T.cell, 64 ), memcpy( d.light_inv_T, scene.light.invT.cell, 64 );
	d.light_size = scene.light.size;
	const float3 lc = scene.GetLightColor(), lp = scene.GetLightPos();
	d.light_color[0] = lc.x, d.light_color[1] = lc.y, d.light_color[2] = lc.z;
	d.light_pos[0] = lp.x, d.light_pos[1] = lp.y, d.light_pos[2] = lp.z;
}

inline void FinishDesc( FlattenedScene& f )
{
	rt_scene_desc& d = f.desc;
	d.blas = f.blas.data(), d.blas_count = (uint32_t)f.blas.size();
	d.obj_material = f.objMaterial.data(), d.obj_count = (uint32_t)f.objMaterial.size();
	d.materials = f.materials.data(), d.material_count = (uint32_t)f.materials.size();
	d.textures = f.textures.data(), d.texture_count = (uint32_t)f.textures.size();
}

#ifdef USE_BVH
// FileScene: one flat SAH BVH over every triangle (file_scene.cpp:4-62); the hit takes Tri::objIdx
inline void Flatten( Tmpl8::FileScene& scene, FlattenedScene& f )
{
	f.desc.kind = RT_SCENE_FLAT;
	rt_blas_desc b = {};
	b.nodes = (const rt_bvh_node*)scene.acc.bvhNodes.data(), b.node_count = scene.acc.nodesUsed;
	b.tris = (const rt_tri*)scene.acc.triangles.data(), b.tri_count = (uint32_t)scene.acc.triangles.size();
	b.tri_indices = (const uint32_t*)scene.acc.triangleIndices.data();
	const mat4 I;
	memcpy( b.T, I.cell, 64 ), memcpy( b.inv_T, I.cell, 64 );
	b.obj_idx = -1, b.mat_idx = -1;
	f.blas.push_back( b );
	for (auto* m : scene.models) f.objMaterial.push_back( m->matIdx );
	FlattenCommon( scene, f );
	FinishDesc( f );
}
#endif

#if defined(USE_KDTree) || defined(USE_Grid)
// the one triangle array KDTree / Grid index into (kdtree.h:33, grid.h:33); nodes / tri_indices stay null
template <class Acc> inline void FlattenTriangles( Tmpl8::FileScene& scene, Acc& acc, FlattenedScene& f )
{
	rt_blas_desc b = {};
	b.tris = (const rt_tri*)acc.triangles.data(), b.tri_count = (uint32_t)acc.triangles.size();
	const mat4 I;
	memcpy( b.T, I.cell, 64 ), memcpy( b.inv_T, I.cell, 64 );
	b.obj_idx = -1, b.mat_idx = -1;
	f.blas.push_back( b );
	for (auto* m : scene.models) f.objMaterial.push_back( m->matIdx );
}
#endif

#if defined(USE_KDTree) || defined(TLAS_USE_KDTree)
// The pointer-linked KDTreeNode graph (blas_kdtree.h:16-25) appended to kdNodes as rt_kd_node[]: the tree's first node is
// its root, children are numbered when their parent is visited (depth first, left first), per-leaf index vectors are
// concatenated; child indices and tri_start are relative to the tree's own first node / first index
inline void FlattenKdTree( const Tmpl8::KDTreeNode* root, std::vector<rt_kd_node>& kdNodes, std::vector<uint32_t>& triIdx )
{
	const size_t nodeBase = kdNodes.size(), idxBase = triIdx.size();
	std::vector<std::pair<const Tmpl8::KDTreeNode*, size_t>> todo;
	kdNodes.push_back( rt_kd_node() );
	todo.push_back( { root, nodeBase } );
	while (!todo.empty())
	{
		const Tmpl8::KDTreeNode* n = todo.back().first;
		const size_t slot = todo.back().second;
		todo.pop_back();
		rt_kd_node k = {};
		k.aabb_min[0] = n->aabbMin.x, k.aabb_min[1] = n->aabbMin.y, k.aabb_min[2] = n->aabbMin.z;
		k.aabb_max[0] = n->aabbMax.x, k.aabb_max[1] = n->aabbMax.y, k.aabb_max[2] = n->aabbMax.z;
		k.split_axis = n->splitAxis, k.split_distance = n->splitDistance;
		k.left = k.right = -1;
		if (n->isLeaf)
		{
			k.tri_start = (uint32_t)(triIdx.size() - idxBase), k.tri_count = (uint32_t)n->triIndices.size();
			triIdx.insert( triIdx.end(), n->triIndices.begin(), n->triIndices.end() );
		}
		else
		{
			k.left = (int32_t)(kdNodes.size() - nodeBase), k.right = k.left + 1;
			kdNodes.push_back( rt_kd_node() ), kdNodes.push_back( rt_kd_node() );
			todo.push_back( { n->right, nodeBase + k.right } ), todo.push_back( { n->left, nodeBase + k.left } );
		}
		kdNodes[slot] = k;
	}
}
#endif

#ifdef USE_KDTree
// FileScene as shipped
inline void Flatten( Tmpl8::FileScene& scene, FlattenedScene& f )
{
	f.desc.kind = RT_SCENE_FLAT_KDTREE;
	FlattenTriangles( scene, scene.acc, f );
	FlattenKdTree( scene.acc.rootNode, f.kdNodes, f.altTriIdx );
	f.desc.kd_nodes = f.kdNodes.data(), f.desc.kd_node_count = (uint32_t)f.kdNodes.size();
	f.desc.kd_tri_indices = f.altTriIdx.data(), f.desc.kd_tri_index_count = (uint32_t)f.altTriIdx.size();
	FlattenCommon( scene, f );
	FinishDesc( f );
}
#endif

#if defined(USE_Grid) || defined(TLAS_USE_Grid)
// Grid / BLASGrid: per-cell index vectors (blas_grid.h:8-11) concatenated in cell order; cell_start is relative to this
// grid's own first index.  The pointers of `g` are filled in by the caller once the vectors have stopped growing.
template <class G> inline void FlattenGrid( G& grid, rt_grid_desc& g, std::vector<uint32_t>& cellStart, std::vector<uint32_t>& triIdx )
{
	const size_t idxBase = triIdx.size();
	for (int i = 0; i < 3; i++)
		g.resolution[i] = grid.resolution[i], g.cell_size[i] = grid.cellSize[i],
		g.bounds_min[i] = grid.localBounds.bmin[i], g.bounds_max[i] = grid.localBounds.bmax[i];
	for (const Tmpl8::GridCell& c : grid.gridCells)
	{
		cellStart.push_back( (uint32_t)(triIdx.size() - idxBase) );
		for (int t : c.triIndices) triIdx.push_back( (uint32_t)t );
	}
	cellStart.push_back( (uint32_t)(triIdx.size() - idxBase) );
	g.index_count = (uint32_t)(triIdx.size() - idxBase);
}
#endif

#ifdef USE_Grid
// FileScene with the uniform grid: per-cell index vectors (blas_grid.h:8-11) concatenated in cell order
inline void Flatten( Tmpl8::FileScene& scene, FlattenedScene& f )
{
	f.desc.kind = RT_SCENE_FLAT_GRID;
	Tmpl8::Grid& g = scene.acc;
	FlattenTriangles( scene, g, f );
	FlattenGrid( g, f.grid, f.gridCellStart, f.altTriIdx );
	f.grid.cell_start = f.gridCellStart.data(), f.grid.tri_indices = f.altTriIdx.data();
	f.desc.grid = &f.grid;
	FlattenCommon( scene, f );
	FinishDesc( f );
}
#endif

#ifdef TLAS_USE_BVH
// Content hash of one BLASBVH's geometry: vertices / normals / uvs / centroids of every triangle (not Tri::objIdx, which the
// loader stamps per object, blas_bvh.cpp:64-80), the node array and the index array.
inline uint64_t MeshHash( const Tmpl8::BLASBVH& b )
{
	uint64_t h = 1469598103934665603ull;
	auto mix = [&]( const void* p, size_t n ) { const unsigned char* c = (const unsigned char*)p; for (size_t i = 0; i < n; i++) h = (h ^ c[i]) * 1099511628211ull; };
	for (const Tri& t : b.triangles) mix( &t, offsetof( rt_tri, obj_idx ) );
	mix( b.bvhNodes.data(), (size_t)b.nodesUsed * sizeof( rt_bvh_node ) );
	mix( b.triangleIndices.data(), b.triangleIndices.size() * sizeof( uint32_t ) );
	return h;
}
inline bool SameMesh( const Tmpl8::BLASBVH& a, const Tmpl8::BLASBVH& b )
{
	if (a.nodesUsed != b.nodesUsed || a.triangles.size() != b.triangles.size()) return false;
	for (size_t i = 0; i < a.triangles.size(); i++)
		if (memcmp( &a.triangles[i], &b.triangles[i], offsetof( rt_tri, obj_idx ) ) != 0) return false;
	return memcmp( a.bvhNodes.data(), b.bvhNodes.data(), (size_t)a.nodesUsed * sizeof( rt_bvh_node ) ) == 0 &&
	       memcmp( a.triangleIndices.data(), b.triangleIndices.data(), a.triangleIndices.size() * sizeof( uint32_t ) ) == 0;
}

// TLASFileScene: one BLASBVH per <object> (tlas_file_scene.cpp:41-55) under the agglomerative TLAS.
// True instancing (README "known issues"): the reference loads and builds every <object> separately, also
// when several use the same OBJ at the same scale.  Objects whose BLAS turned out IDENTICAL (same triangles, nodes, indices:
// hash, then full compare) are handed over with the arrays of the first of them, and rt_scene_create keeps ONE device copy
// per distinct array set; the instances differ only in (T, invT, objIdx, matIdx).  Hits are unchanged: in a TLAS scene the
// hit's objIdx is the BLAS' own (blas_bvh.cpp:297), not the triangle's.
// deviceTlas: leave TLASBVH::Build to the device (32-bit children: no 32 767-instance cap) instead of taking the host's nodes.
inline void Flatten( Tmpl8::TLASFileScene& scene, FlattenedScene& f, bool deviceTlas = false )
{
	f.desc.kind = RT_SCENE_TLAS;
	std::vector<std::pair<uint64_t, Tmpl8::BLASBVH*>> distinct;
	for (Tmpl8::BLASBVH* blas : scene.tlas.blas)
	{
		const uint64_t h = MeshHash( *blas );
		Tmpl8::BLASBVH* src = blas;
		for (auto& d : distinct)
			if (d.first == h && SameMesh( *d.second, *blas )) { src = d.second; break; }
		if (src == blas) distinct.push_back( { h, blas } );
		rt_blas_desc b = {};
		b.nodes = (const rt_bvh_node*)src->bvhNodes.data(), b.node_count = src->nodesUsed;
		b.tris = (const rt_tri*)src->triangles.data(), b.tri_count = (uint32_t)src->triangles.size();
		b.tri_indices = (const uint32_t*)src->triangleIndices.data();
		memcpy( b.T, blas->T.cell, 64 ), memcpy( b.inv_T, blas->invT.cell, 64 );
		b.obj_idx = blas->objIdx, b.mat_idx = blas->matIdx;
		f.blas.push_back( b );
		f.objMaterial.push_back( blas->matIdx );
	}
	if (!deviceTlas)
	{
		f.desc.tlas_nodes = (const rt_tlas_node*)scene.tlas.tlasNode;
		f.desc.tlas_node_count = scene.tlas.nodesUsed;
	}
	FlattenCommon( scene, f );
	FinishDesc( f );
}
#endif

#if defined(TLAS_USE_KDTree) || defined(TLAS_USE_Grid)
// TLASFileScene over per-object KD-trees / grids: the same agglomerative TLAS (tlas_kdtree.cpp:17-70 = tlas_grid.cpp:17-70),
// BLASKDTree / BLASGrid leaves (needs read access to TLASKDTree::tlasNode / TLASGrid::tlasNode and, for the grid, to
// BLASGrid::resolution / cellSize / triangles / gridCells, private in the reference)
inline void Flatten( Tmpl8::TLASFileScene& scene, FlattenedScene& f )
{
	struct Range { size_t node, nodes, idx, idxs, cell; };
	std::vector<Range> ranges;
#ifdef TLAS_USE_KDTree
	f.desc.kind = RT_SCENE_TLAS_KDTREE;
	static_assert(sizeof( Tmpl8::TLASKDTreeNode ) == sizeof( rt_tlas_node ), "TLASKDTreeNode layout (tlas_kdtree.h:6-13)");
#else
	f.desc.kind = RT_SCENE_TLAS_GRID;
	static_assert(sizeof( Tmpl8::TLASGridNode ) == sizeof( rt_tlas_node ), "TLASGridNode layout (tlas_grid.h:7-14)");
	f.blasGrids.resize( scene.tlas.blas.size() );
#endif
	size_t i = 0;
	for (auto* blas : scene.tlas.blas)
	{
		rt_blas_desc b = {};
		b.tris = (const rt_tri*)blas->triangles.data(), b.tri_count = (uint32_t)blas->triangles.size();
		memcpy( b.T, blas->T.cell, 64 ), memcpy( b.inv_T, blas->invT.cell, 64 );
		b.obj_idx = blas->objIdx, b.mat_idx = blas->matIdx;
		f.blas.push_back( b );
		f.objMaterial.push_back( blas->matIdx );
		Range r = { f.kdNodes.size(), 0, f.altTriIdx.size(), 0, f.gridCellStart.size() };
#ifdef TLAS_USE_KDTree
		FlattenKdTree( blas->rootNode, f.kdNodes, f.altTriIdx );
#else
		FlattenGrid( *blas, f.blasGrids[i], f.gridCellStart, f.altTriIdx );
#endif
		r.nodes = f.kdNodes.size() - r.node, r.idxs = f.altTriIdx.size() - r.idx;
		ranges.push_back( r );
		i++;
	}
	f.blasAccel.resize( ranges.size() );
	for (i = 0; i < ranges.size(); i++) // pointers only now: the vectors no longer grow
	{
		rt_blas_accel a = {};
#ifdef TLAS_USE_KDTree
		a.kd_nodes = f.kdNodes.data() + ranges[i].node, a.kd_node_count = (uint32_t)ranges[i].nodes;
		a.kd_tri_indices = f.altTriIdx.data() + ranges[i].idx, a.kd_tri_index_count = (uint32_t)ranges[i].idxs;
#else
		f.blasGrids[i].cell_start = f.gridCellStart.data() + ranges[i].cell;
		f.blasGrids[i].tri_indices = f.altTriIdx.data() + ranges[i].idx;
		a.grid = &f.blasGrids[i];
#endif
		f.blasAccel[i] = a;
	}
	f.desc.blas_accel = f.blasAccel.data();
	f.desc.tlas_nodes = (const rt_tlas_node*)scene.tlas.tlasNode;
	f.desc.tlas_node_count = scene.tlas.nodesUsed;
	FlattenCommon( scene, f );
	FinishDesc( f );
}
#endif

// ------------------------------------------------------------------------------------------------
template <class HostScene> class GpuScene : public Tmpl8::BaseScene
{
public:
	explicit GpuScene( const std::string& filePath, int device = 0 ) : host( filePath )
	{
		Flatten( host, flat );
		check( rt_scene_create( &flat.desc, device, 0, &dev ), "rt_scene_create" );
	}
	GpuScene( const GpuScene& ) = delete;
	GpuScene& operator=( const GpuScene& ) = delete;
	~GpuScene() { rt_scene_destroy( dev ); }
	// BaseScene, answered by the host scene (immutable after construction)
	void SetTime( float t ) override { host.SetTime( t ); }
	float3 GetSkyColor( const Tmpl8::Ray& ray ) const override { return host.GetSkyColor( ray ); }
	float3 GetLightPos() const override { return host.GetLightPos(); }
	float3 GetLightColor() const override { return host.GetLightColor(); }
	float3 GetAlbedo( int objIdx, float3 I ) const override { return host.GetAlbedo( objIdx, I ); }
	HitInfo GetHitInfo( const Tmpl8::Ray& ray, const float3 I ) override { return host.GetHitInfo( ray, I ); }
	int GetTriangleCount() const override { return host.GetTriangleCount(); }
	// BaseScene, answered by the GPU.  Per-ray, synchronous, in-out by reference like the reference
	// (bvh.cpp:220 writes t / objIdx / triIdx / barycentric; a miss leaves objIdx == -1, ray.h:36).
	void FindNearest( Tmpl8::Ray& ray ) override { FindNearest( &ray, 1 ); }
	bool IsOccluded( const Tmpl8::Ray& ray ) override
	{
		uint8_t o = 0;
		IsOccluded( &ray, 1, &o );
		return o != 0;
	}
	// batched forms: what a caller that owns many rays should use
	void FindNearest( Tmpl8::Ray* rays, size_t n )
	{
		in.resize( n ), out.resize( n );
		for (size_t i = 0; i < n; i++) Pack( rays[i], in[i] );
		check( rt_find_nearest( dev, in.data(), out.data(), n ), "rt_find_nearest" );
		for (size_t i = 0; i < n; i++)
		{
			Tmpl8::Ray& r = rays[i];
			r.t = out[i].t, r.barycentric = float2( out[i].u, out[i].v );
			r.objIdx = out[i].obj_idx, r.triIdx = out[i].tri_idx;
			r.traversed = out[i].traversed, r.tested = out[i].tested;
		}
	}
	void IsOccluded( const Tmpl8::Ray* rays, size_t n, uint8_t* occluded )
	{
		in.resize( n );
		for (size_t i = 0; i < n; i++) Pack( rays[i], in[i] );
		check( rt_is_occluded( dev, in.data(), occluded, n ), "rt_is_occluded" );
	}
	rt_scene* Handle() const { return dev; }
	// the tables handed to rt_scene_create (they point into `host`'s containers): a multi-GPU renderer replicates the scene from them
	const rt_scene_desc& Desc() const { return flat.desc; }
	rt_scene_info Info() const { rt_scene_info i = {}; check( rt_scene_get_info( dev, &i ), "rt_scene_get_info" ); return i; }
public:
	HostScene host;
private:
	FlattenedScene flat;
	static void Pack( const Tmpl8::Ray& r, rt_ray& p )
	{
		p.O[0] = r.O.x, p.O[1] = r.O.y, p.O[2] = r.O.z, p.tmax = r.t;
		p.D[0] = r.D.x, p.D[1] = r.D.y, p.D[2] = r.D.z, p.inside = r.inside ? 1 : 0;
	}
	rt_scene* dev = nullptr;
	std::vector<rt_ray> in;
	std::vector<rt_hit> out;
};

// ------------------------------------------------------------------------------------------------
template <class HostScene, int INTEGRATOR> class GpuRenderer : public TheApp
{
public:
	// devices: the GPUs that render.  One device = rt_renderer on it.  Several (path tracer) = rt_multi_renderer: the tile jobs
	// of Renderer::Tick (3. PathTracer/renderer.cpp:144-168) are dealt to the GPUs, tile k to device k mod n, and every GPU
	// writes its tiles into ONE accumulator on devices[0] through peer-mapped memory - the image is bit-identical to one GPU's.
	explicit GpuRenderer( const std::string& scenePath, int device = 0 ) : scene( scenePath, device ), devices( 1, device ) {}
	GpuRenderer( const std::string& scenePath, const std::vector<int>& gpus ) : scene( scenePath, gpus.empty() ? 0 : gpus[0] ), devices( gpus.empty() ? std::vector<int>( 1, 0 ) : gpus ) {}
	~GpuRenderer()
	{
		if (dev) rt_renderer_destroy( dev );
		if (multi) rt_multi_renderer_destroy( multi );
		if (accumulator) FREE64( accumulator );
	}
	// Renderer::Init (renderer.cpp:8-13)
	void Init() override
	{
		accumulator = (float4*)MALLOC64( (size_t)SCRWIDTH * SCRHEIGHT * 16 );
		memset( accumulator, 0, (size_t)SCRWIDTH * SCRHEIGHT * 16 );
		rt_render_params p;
		rt_render_params_default( &p, INTEGRATOR, SCRWIDTH, SCRHEIGHT );
		p.depth_limit = depthLimit, p.epsilon = EPSILON;
		// one Tick per frame is how the reference runs: render `lookahead` frames per launch, reveal one per Tick
		p.lookahead_frames = INTEGRATOR == RT_INTEGRATOR_PATH ? lookahead : 0;
		if (INTEGRATOR == RT_INTEGRATOR_PATH && devices.size() > 1)
			check( rt_multi_renderer_create( &scene.Desc(), 0, devices.data(), (int)devices.size(), &p, &multi ), "rt_multi_renderer_create" );
		else check( rt_renderer_create( scene.Handle(), &p, &dev ), "rt_renderer_create" );
		createdDepthLimit = depthLimit;
	}
	// Renderer::ClearAccumulator (3. PathTracer/renderer.cpp:15-18)
	void ClearAccumulator()
	{
		memset( accumulator, 0, (size_t)SCRWIDTH * SCRHEIGHT * 16 );
		check( multi ? rt_multi_renderer_clear( multi ) : rt_renderer_clear( dev ), "rt_renderer_clear" );
	}
	// Renderer::Tick (3. PathTracer/renderer.cpp:144-168, 2. WhittedStyle/renderer.cpp:131-190)
	void Tick( float deltaTime ) override
	{
		if (depthLimit != createdDepthLimit) // the UI may change depthLimit between frames
		{
			if (dev) rt_renderer_destroy( dev ), dev = nullptr;
			if (multi) rt_multi_renderer_destroy( multi ), multi = nullptr;
			FREE64( accumulator );
			Init();
		}
		if (animating) scene.SetTime( anim_time += deltaTime * 0.002f ), ClearAccumulator();
		Submit( 1 );
		const float scale = INTEGRATOR == RT_INTEGRATOR_PATH ? 1.0f / (spp + passes) : 1.0f; // renderer.cpp:119
		if (screen) check( multi ? rt_multi_renderer_read_pixels( multi, scale, (uint32_t*)screen->pixels ) : rt_renderer_read_pixels( dev, scale, (uint32_t*)screen->pixels ), "rt_renderer_read_pixels" );
		if (INTEGRATOR == RT_INTEGRATOR_PATH)
		{
			if (camera.HandleInput( deltaTime )) ClearAccumulator();
			else spp += passes;
		}
		else camera.HandleInput( deltaTime );
	}
	// `frames` Ticks in one call: all (tile, frame) RNG streams are in flight together on the GPU(s)
	void Render( int frames )
	{
		Submit( frames );
		if (INTEGRATOR == RT_INTEGRATOR_PATH) spp += frames * passes;
	}
	rt_renderer* Handle() const { return dev; }
	rt_multi_renderer* MultiHandle() const { return multi; }
	int DeviceCount() const { return (int)devices.size(); }
	// data members, as in the reference's Renderer
	int2 mousePos;
	float4* accumulator = nullptr;
	GpuScene<HostScene> scene;
	Tmpl8::Camera camera;
	int spp = 1, passes = 1;
	bool animating = false;
	float energy = 0, anim_time = 0;
	int depthLimit = 5;
	int lookahead = 32; // frames rendered ahead of the Tick sequence (set before Init; 0 = off)
private:
	static void Store( float* p, const float3& v ) { p[0] = v.x, p[1] = v.y, p[2] = v.z; }
	// camera + passes + `frames` frames from spp + accumulator back to the host copy the reference's callers read
	void Submit( int frames )
	{
		rt_camera c;
		Store( c.pos, camera.camPos ), Store( c.top_left, camera.topLeft );
		Store( c.top_right, camera.topRight ), Store( c.bottom_left, camera.bottomLeft );
		if (multi)
		{
			// the UI's "spp" slider changes `passes` between frames (renderer.cpp:182): samples per pixel per Tick
			check( rt_multi_renderer_set_passes( multi, passes ), "rt_multi_renderer_set_passes" );
			check( rt_multi_renderer_set_camera( multi, &c ), "rt_multi_renderer_set_camera" );
			check( rt_multi_renderer_render( multi, spp, frames, passes ), "rt_multi_renderer_render" );
			check( rt_multi_renderer_read_accumulator( multi, (float*)accumulator ), "rt_multi_renderer_read_accumulator" );
			return;
		}
		if (INTEGRATOR == RT_INTEGRATOR_PATH) check( rt_renderer_set_passes( dev, passes ), "rt_renderer_set_passes" );
		check( rt_renderer_set_camera( dev, &c ), "rt_renderer_set_camera" );
		check( rt_renderer_render( dev, spp, frames, passes ), "rt_renderer_render" );
		check( rt_renderer_read_accumulator( dev, (float*)accumulator ), "rt_renderer_read_accumulator" );
	}
	std::vector<int> devices;
	rt_renderer* dev = nullptr;
	rt_multi_renderer* multi = nullptr;
	int createdDepthLimit = -1;
};

} // namespace rtb200
