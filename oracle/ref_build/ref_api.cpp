// C API over the reference's own classes (TEST INFRASTRUCTURE ONLY; compiled by build_ref.py as the
// last file of a unity translation unit that contains the reference's unmodified hot-path sources).
//
// This is OUR driver code: it instantiates the reference's Renderer (which owns the scene and the
// camera), calls its public methods, and flattens the scene's public containers into the .rtscene
// chunk file that tests/ feed to both the CUDA library and the C port oracle.  It plays the role of
// template/template.cpp's frame loop (template.cpp:141,168,305-338) without a window.
#include <unistd.h>

int g_ref_scrwidth = 1024, g_ref_scrheight = 640; // template/camera.h:4-5 defaults
std::string g_ref_scene_path;

static Renderer* g_renderer = nullptr;
static double g_last_tick_seconds = 0;

static inline void api_ray( Ray& r, const float* O, const float* D, float tmax )
{
	r = Ray( float3( O[0], O[1], O[2] ), float3( D[0], D[1], D[2] ), tmax );
}

extern "C" {

// workdir must be a directory whose parent holds `assets/` (scene files use "../assets/...").
int ref_create( const char* scene_xml, const char* workdir, int width, int height )
{
	try
	{
		if (workdir && workdir[0] && chdir( workdir ) != 0) return -2;
		g_ref_scrwidth = width, g_ref_scrheight = height;
		g_ref_scene_path = scene_xml;
		g_renderer = new Renderer();
		g_renderer->screen = new Surface( width, height );
		g_renderer->Init();
		return 0;
	}
	catch (const std::exception& e)
	{
		fprintf( stderr, "ref_create: %s\n", e.what() );
		return -1;
	}
}

int ref_width() { return g_ref_scrwidth; }
int ref_height() { return g_ref_scrheight; }
int ref_threads() { return omp_get_max_threads(); }

void ref_set_camera( const float* pos, const float* target )
{
	g_renderer->camera.SetCameraState( float3( pos[0], pos[1], pos[2] ), float3( target[0], target[1], target[2] ) );
}

// camPos, topLeft, topRight, bottomLeft (12 floats)
void ref_get_camera( float* out )
{
	Camera& c = g_renderer->camera;
	const float3 v[4] = { c.camPos, c.topLeft, c.topRight, c.bottomLeft };
	for (int i = 0; i < 4; i++) out[3 * i] = v[i].x, out[3 * i + 1] = v[i].y, out[3 * i + 2] = v[i].z;
}

void ref_set_depth_limit( int d ) { g_renderer->depthLimit = d; }

// Scene::FindNearest over a batch of rays (SoA in, SoA out)
void ref_find_nearest( int n, const float* O, const float* D, const float* tmax,
	float* t, float* u, float* v, int* objIdx, int* triIdx, int* traversed, int* tested )
{
#pragma omp parallel for schedule( dynamic, 256 )
	for (int i = 0; i < n; i++)
	{
		Ray r;
		api_ray( r, O + 3 * i, D + 3 * i, tmax ? tmax[i] : 1e34f );
		g_renderer->scene.FindNearest( r );
		t[i] = r.t, u[i] = r.barycentric.x, v[i] = r.barycentric.y, objIdx[i] = r.objIdx, triIdx[i] = r.triIdx;
		if (traversed) traversed[i] = r.traversed;
		if (tested) tested[i] = r.tested;
	}
}

void ref_is_occluded( int n, const float* O, const float* D, const float* tmax, unsigned char* out )
{
#pragma omp parallel for schedule( dynamic, 256 )
	for (int i = 0; i < n; i++)
	{
		Ray r;
		api_ray( r, O + 3 * i, D + 3 * i, tmax[i] );
		out[i] = g_renderer->scene.IsOccluded( r ) ? 1 : 0;
	}
}

// camera.GetPrimaryRay(x, y) for every pixel (no jitter) + FindNearest; also returns the rays
void ref_primary_hits( float* O, float* D, float* t, float* u, float* v, int* objIdx, int* triIdx, int* traversed, int* tested )
{
	const int W = g_ref_scrwidth, H = g_ref_scrheight;
#pragma omp parallel for schedule( dynamic, 4 )
	for (int y = 0; y < H; y++) for (int x = 0; x < W; x++)
	{
		const int i = x + y * W;
		Ray r = g_renderer->camera.GetPrimaryRay( (float)x, (float)y );
		if (O) O[3 * i] = r.O.x, O[3 * i + 1] = r.O.y, O[3 * i + 2] = r.O.z;
		if (D) D[3 * i] = r.D.x, D[3 * i + 1] = r.D.y, D[3 * i + 2] = r.D.z;
		g_renderer->scene.FindNearest( r );
		t[i] = r.t, u[i] = r.barycentric.x, v[i] = r.barycentric.y, objIdx[i] = r.objIdx, triIdx[i] = r.triIdx;
		if (traversed) traversed[i] = r.traversed;
		if (tested) tested[i] = r.tested;
	}
}

// shading queries for a batch of hits: GetHitInfo normal/uv + material albedo, GetSkyColor for misses
void ref_hit_info( int n, const float* O, const float* D, const float* t, const float* u, const float* v,
	const int* objIdx, const int* triIdx, float* N, float* uv, float* albedo )
{
	for (int i = 0; i < n; i++)
	{
		Ray r;
		api_ray( r, O + 3 * i, D + 3 * i, t[i] );
		r.barycentric = float2( u[i], v[i] ), r.objIdx = objIdx[i], r.triIdx = triIdx[i];
		float3 a;
		if (r.objIdx == -1)
		{
			a = g_renderer->scene.GetSkyColor( r );
			N[3 * i] = N[3 * i + 1] = N[3 * i + 2] = 0, uv[2 * i] = uv[2 * i + 1] = 0;
		}
		else
		{
			const float3 I = r.O + r.t * r.D;
			HitInfo h = g_renderer->scene.GetHitInfo( r, I );
			N[3 * i] = h.normal.x, N[3 * i + 1] = h.normal.y, N[3 * i + 2] = h.normal.z;
			uv[2 * i] = h.uv.x, uv[2 * i + 1] = h.uv.y;
			a = h.material->GetAlbedo( h.uv );
		}
		albedo[3 * i] = a.x, albedo[3 * i + 1] = a.y, albedo[3 * i + 2] = a.z;
	}
}

// `frames` calls of Renderer::Tick (one frame = one sample per pixel for the path tracer)
void ref_tick( int frames )
{
	Timer t;
	for (int i = 0; i < frames; i++) g_renderer->Tick( 0 );
	g_last_tick_seconds = t.elapsed();
}
double ref_last_tick_seconds() { return g_last_tick_seconds; }
const float* ref_accumulator() { return (const float*)g_renderer->accumulator; }
const unsigned int* ref_screen() { return g_renderer->screen->pixels; }

#ifdef REF_INTEGRATOR_PT
int ref_spp() { return g_renderer->spp; }
void ref_reset( int spp ) { g_renderer->ClearAccumulator(); g_renderer->spp = spp; }
void ref_set_passes( int p ) { g_renderer->passes = p; }
float ref_energy() { return g_renderer->energy; }
#else
int ref_spp() { return 1; }
void ref_reset( int ) { memset( g_renderer->accumulator, 0, (size_t)g_ref_scrwidth * g_ref_scrheight * 16 ); }
void ref_set_passes( int ) {}
float ref_energy() { return 0; }
#endif

// ---------------------------------------------------------------------------------------------
// flattening: scene -> .rtscene chunk file (format documented in cpu-ray-tracer_b200/scene_file.py)
// ---------------------------------------------------------------------------------------------
struct ChunkWriter
{
	FILE* f = nullptr;
	uint32_t count = 0;
	bool open( const char* path )
	{
		f = fopen( path, "wb" );
		if (!f) return false;
		fwrite( "RTSCN001", 1, 8, f );
		fwrite( &count, 4, 1, f ); // patched on close
		uint32_t pad = 0;
		fwrite( &pad, 4, 1, f );
		return true;
	}
	void chunk( const char* name, const void* data, uint64_t nbytes )
	{
		char nm[24] = {};
		strncpy( nm, name, 23 );
		fwrite( nm, 1, 24, f );
		fwrite( &nbytes, 8, 1, f );
		if (nbytes) fwrite( data, 1, nbytes, f );
		const uint64_t padded = (nbytes + 7) & ~7ull;
		const char zero[8] = {};
		if (padded > nbytes) fwrite( zero, 1, padded - nbytes, f );
		count++;
	}
	void close()
	{
		fseek( f, 8, SEEK_SET );
		fwrite( &count, 4, 1, f );
		fclose( f );
	}
};

struct FlatBlas // mirrors rt_blas_desc scalars (include/rt_b200.h)
{
	uint32_t node_offset, node_count, tri_offset, tri_count;
	float T[16], invT[16];
	int32_t obj_idx, mat_idx;
};
struct FlatMaterial // mirrors rt_material
{
	float reflectivity, refractivity, absorption[3], albedo[3];
	int32_t is_light, texture;
};
struct FlatTexture { uint64_t pixel_offset; int32_t width, height; };
struct FlatHeader
{
	int32_t kind; // 0 = FileScene flat BVH, 1 = TLASFileScene
	int32_t skydome_texture, floor_texture, reserved;
	float floor_n[3], floor_d, floor_invto;
	float light_T[16], light_invT[16], light_size;
	float light_color[3], light_pos[3];
};

static int add_texture( const Texture* tex, std::vector<FlatTexture>& table, std::vector<uint>& pixels )
{
	if (!tex || tex->width == 0) return -1;
	FlatTexture t;
	t.pixel_offset = pixels.size(), t.width = tex->width, t.height = tex->height;
	pixels.insert( pixels.end(), tex->pixels.begin(), tex->pixels.end() );
	table.push_back( t );
	return (int)table.size() - 1;
}

struct FlatKdNode { float aabbMin[3]; int32_t left; float aabbMax[3]; int32_t right; int32_t splitAxis; float splitDistance; uint32_t triStart, triCount; };
struct FlatGridHeader { int32_t resolution[3]; float cellSize[3], boundsMin[3], boundsMax[3]; };
struct FlatBlasKd { uint32_t node_offset, node_count, idx_offset, idx_count; };                 // per BLAS, chunk "blas_kd_table"
struct FlatBlasGrid { FlatGridHeader h; uint32_t cell_offset, cell_count, idx_offset, idx_count; }; // per BLAS, chunk "blas_grid_table"

#if defined(USE_KDTree) || defined(TLAS_USE_KDTree)
// KDTreeNode graph (blas_kdtree.h:16-25) -> flat nodes: root first, children numbered when their parent is visited (depth
// first, left first), child indices and triStart relative to this tree's first node / first index
static void flatten_kd( KDTreeNode* root, std::vector<FlatKdNode>& kdNodes, std::vector<uint>& kdTriIdx )
{
	const size_t nodeBase = kdNodes.size(), idxBase = kdTriIdx.size();
	std::vector<std::pair<KDTreeNode*, size_t>> todo; // node, slot
	kdNodes.push_back( FlatKdNode() );
	todo.push_back( { root, nodeBase } );
	while (!todo.empty())
	{
		auto it = todo.back();
		todo.pop_back();
		KDTreeNode* n = it.first;
		FlatKdNode f = {};
		f.aabbMin[0] = n->aabbMin.x, f.aabbMin[1] = n->aabbMin.y, f.aabbMin[2] = n->aabbMin.z;
		f.aabbMax[0] = n->aabbMax.x, f.aabbMax[1] = n->aabbMax.y, f.aabbMax[2] = n->aabbMax.z;
		f.splitAxis = n->splitAxis, f.splitDistance = n->splitDistance;
		f.left = f.right = -1;
		if (n->isLeaf)
		{
			f.triStart = (uint32_t)(kdTriIdx.size() - idxBase), f.triCount = (uint32_t)n->triIndices.size();
			kdTriIdx.insert( kdTriIdx.end(), n->triIndices.begin(), n->triIndices.end() );
		}
		else
		{
			f.left = (int32_t)(kdNodes.size() - nodeBase), f.right = f.left + 1;
			kdNodes.push_back( FlatKdNode() ), kdNodes.push_back( FlatKdNode() );
			todo.push_back( { n->right, nodeBase + f.right } ), todo.push_back( { n->left, nodeBase + f.left } );
		}
		kdNodes[it.second] = f;
	}
}
#endif

#if defined(USE_Grid) || defined(TLAS_USE_Grid)
} // extern "C" (templates need C++ linkage)
template <class G> static void flatten_grid( G& g, FlatGridHeader& gh, std::vector<uint>& cellStart, std::vector<uint>& gridTriIdx )
{
	const size_t idxBase = gridTriIdx.size();
	for (int i = 0; i < 3; i++)
		gh.resolution[i] = g.resolution[i], gh.cellSize[i] = g.cellSize[i], gh.boundsMin[i] = g.localBounds.bmin[i], gh.boundsMax[i] = g.localBounds.bmax[i];
	for (const GridCell& c : g.gridCells)
	{
		cellStart.push_back( (uint)(gridTriIdx.size() - idxBase) );
		for (int t : c.triIndices) gridTriIdx.push_back( (uint)t );
	}
	cellStart.push_back( (uint)(gridTriIdx.size() - idxBase) );
}
extern "C" {
#endif

int ref_flatten( const char* out_path )
{
	auto& scene = g_renderer->scene;
	ChunkWriter w;
	if (!w.open( out_path )) return -1;
	std::vector<FlatBlas> blasTable;
	std::vector<BVHNode> nodes;
	std::vector<Tri> tris;
	std::vector<uint> triIdx;
	std::vector<int32_t> objMaterial;
	FlatHeader h = {};
#if defined(REF_SCENE_FILE) && defined(USE_KDTree)
	// FileScene as shipped (file_scene.h:10-12): spatial-median KD-tree, pointer nodes, per-leaf index vectors
	// (kdtree.cpp:45-112).  Flattened depth first: node 0 = root, children by index, leaf lists concatenated.
	h.kind = 2;
	std::vector<FlatKdNode> kdNodes;
	std::vector<uint> kdTriIdx;
	{
		flatten_kd( scene.acc.rootNode, kdNodes, kdTriIdx );
		FlatBlas b = {};
		b.tri_count = (uint32_t)scene.acc.triangles.size();
		mat4 I;
		memcpy( b.T, I.cell, 64 ), memcpy( b.invT, I.cell, 64 );
		b.obj_idx = -1, b.mat_idx = -1;
		blasTable.push_back( b );
		tris = scene.acc.triangles;
		for (auto* m : scene.models) objMaterial.push_back( m->matIdx );
	}
	w.chunk( "kd_nodes", kdNodes.data(), kdNodes.size() * sizeof( FlatKdNode ) );
	w.chunk( "kd_tri_indices", kdTriIdx.data(), kdTriIdx.size() * sizeof( uint ) );
#elif defined(REF_SCENE_FILE) && defined(USE_Grid)
	// FileScene with the uniform grid (grid.cpp:4-60): per-cell index vectors concatenated in cell order
	h.kind = 3;
	FlatGridHeader gh = {};
	std::vector<uint> cellStart, gridTriIdx;
	{
		Grid& g = scene.acc;
		flatten_grid( g, gh, cellStart, gridTriIdx );
		FlatBlas b = {};
		b.tri_count = (uint32_t)g.triangles.size();
		mat4 I;
		memcpy( b.T, I.cell, 64 ), memcpy( b.invT, I.cell, 64 );
		b.obj_idx = -1, b.mat_idx = -1;
		blasTable.push_back( b );
		tris = g.triangles;
		for (auto* m : scene.models) objMaterial.push_back( m->matIdx );
	}
	w.chunk( "grid_header", &gh, sizeof( gh ) );
	w.chunk( "grid_cell_start", cellStart.data(), cellStart.size() * sizeof( uint ) );
	w.chunk( "grid_tri_indices", gridTriIdx.data(), gridTriIdx.size() * sizeof( uint ) );
#elif defined(REF_SCENE_FILE)
	h.kind = 0;
	{
		FlatBlas b = {};
		b.node_offset = 0, b.node_count = scene.acc.nodesUsed, b.tri_offset = 0, b.tri_count = (uint32_t)scene.acc.triangles.size();
		mat4 I;
		memcpy( b.T, I.cell, 64 ), memcpy( b.invT, I.cell, 64 );
		b.obj_idx = -1, b.mat_idx = -1;
		blasTable.push_back( b );
		nodes.assign( scene.acc.bvhNodes.begin(), scene.acc.bvhNodes.begin() + scene.acc.nodesUsed );
		tris = scene.acc.triangles;
		triIdx = scene.acc.triangleIndices;
		for (auto* m : scene.models) objMaterial.push_back( m->matIdx );
	}
#elif defined(TLAS_USE_KDTree)
	// TLASFileScene over per-object KD-trees: the same agglomerative TLAS (tlas_kdtree.cpp:17-70), BLASKDTree leaves
	h.kind = 4;
	std::vector<FlatKdNode> kdNodes;
	std::vector<uint> kdTriIdx;
	std::vector<FlatBlasKd> kdTable;
	for (BLASKDTree* blas : scene.tlas.blas)
	{
		FlatBlas b = {};
		b.tri_offset = (uint32_t)tris.size(), b.tri_count = (uint32_t)blas->triangles.size();
		memcpy( b.T, blas->T.cell, 64 ), memcpy( b.invT, blas->invT.cell, 64 );
		b.obj_idx = blas->objIdx, b.mat_idx = blas->matIdx;
		blasTable.push_back( b );
		FlatBlasKd k = { (uint32_t)kdNodes.size(), 0, (uint32_t)kdTriIdx.size(), 0 };
		flatten_kd( blas->rootNode, kdNodes, kdTriIdx );
		k.node_count = (uint32_t)kdNodes.size() - k.node_offset, k.idx_count = (uint32_t)kdTriIdx.size() - k.idx_offset;
		kdTable.push_back( k );
		tris.insert( tris.end(), blas->triangles.begin(), blas->triangles.end() );
		objMaterial.push_back( blas->matIdx );
	}
	w.chunk( "tlas_nodes", scene.tlas.tlasNode, sizeof( TLASKDTreeNode ) * scene.tlas.nodesUsed );
	w.chunk( "blas_kd_table", kdTable.data(), kdTable.size() * sizeof( FlatBlasKd ) );
	w.chunk( "kd_nodes", kdNodes.data(), kdNodes.size() * sizeof( FlatKdNode ) );
	w.chunk( "kd_tri_indices", kdTriIdx.data(), kdTriIdx.size() * sizeof( uint ) );
#elif defined(TLAS_USE_Grid)
	// TLASFileScene over per-object uniform grids (tlas_grid.cpp:17-70, blas_grid.cpp)
	h.kind = 5;
	std::vector<uint> cellStart, gridTriIdx;
	std::vector<FlatBlasGrid> gridTable;
	for (BLASGrid* blas : scene.tlas.blas)
	{
		FlatBlas b = {};
		b.tri_offset = (uint32_t)tris.size(), b.tri_count = (uint32_t)blas->triangles.size();
		memcpy( b.T, blas->T.cell, 64 ), memcpy( b.invT, blas->invT.cell, 64 );
		b.obj_idx = blas->objIdx, b.mat_idx = blas->matIdx;
		blasTable.push_back( b );
		FlatBlasGrid g = {};
		g.cell_offset = (uint32_t)cellStart.size(), g.idx_offset = (uint32_t)gridTriIdx.size();
		flatten_grid( *blas, g.h, cellStart, gridTriIdx );
		g.cell_count = (uint32_t)cellStart.size() - g.cell_offset - 1, g.idx_count = (uint32_t)gridTriIdx.size() - g.idx_offset;
		gridTable.push_back( g );
		tris.insert( tris.end(), blas->triangles.begin(), blas->triangles.end() );
		objMaterial.push_back( blas->matIdx );
	}
	w.chunk( "tlas_nodes", scene.tlas.tlasNode, sizeof( TLASGridNode ) * scene.tlas.nodesUsed );
	w.chunk( "blas_grid_table", gridTable.data(), gridTable.size() * sizeof( FlatBlasGrid ) );
	w.chunk( "grid_cell_start", cellStart.data(), cellStart.size() * sizeof( uint ) );
	w.chunk( "grid_tri_indices", gridTriIdx.data(), gridTriIdx.size() * sizeof( uint ) );
#else
	h.kind = 1;
	for (BLASBVH* blas : scene.tlas.blas)
	{
		FlatBlas b = {};
		b.node_offset = (uint32_t)nodes.size(), b.node_count = blas->nodesUsed;
		b.tri_offset = (uint32_t)tris.size(), b.tri_count = (uint32_t)blas->triangles.size();
		memcpy( b.T, blas->T.cell, 64 ), memcpy( b.invT, blas->invT.cell, 64 );
		b.obj_idx = blas->objIdx, b.mat_idx = blas->matIdx;
		blasTable.push_back( b );
		nodes.insert( nodes.end(), blas->bvhNodes.begin(), blas->bvhNodes.begin() + blas->nodesUsed );
		tris.insert( tris.end(), blas->triangles.begin(), blas->triangles.end() );
		triIdx.insert( triIdx.end(), blas->triangleIndices.begin(), blas->triangleIndices.end() );
		objMaterial.push_back( blas->matIdx );
	}
	w.chunk( "tlas_nodes", scene.tlas.tlasNode, sizeof( TLASBVHNode ) * scene.tlas.nodesUsed );
#endif
	std::vector<FlatTexture> texTable;
	std::vector<uint> texPixels;
	h.skydome_texture = add_texture( &scene.skydome, texTable, texPixels );
	h.floor_texture = add_texture( scene.primitiveMaterials[1].textureDiffuse.get(), texTable, texPixels );
	std::vector<FlatMaterial> mats;
	for (Material* m : scene.materials)
	{
		FlatMaterial fm = {};
		fm.reflectivity = m->reflectivity, fm.refractivity = m->refractivity;
		fm.absorption[0] = m->absorption.x, fm.absorption[1] = m->absorption.y, fm.absorption[2] = m->absorption.z;
		fm.albedo[0] = m->albedo.x, fm.albedo[1] = m->albedo.y, fm.albedo[2] = m->albedo.z;
		fm.is_light = m->isLight ? 1 : 0;
		fm.texture = add_texture( m->textureDiffuse.get(), texTable, texPixels );
		mats.push_back( fm );
	}
	h.floor_n[0] = scene.floor.N.x, h.floor_n[1] = scene.floor.N.y, h.floor_n[2] = scene.floor.N.z;
	h.floor_d = scene.floor.d, h.floor_invto = scene.floor.invto;
	memcpy( h.light_T, scene.light.T.cell, 64 ), memcpy( h.light_invT, scene.light.invT.cell, 64 );
	h.light_size = scene.light.size;
	const float3 lc = scene.GetLightColor(), lp = scene.GetLightPos();
	h.light_color[0] = lc.x, h.light_color[1] = lc.y, h.light_color[2] = lc.z;
	h.light_pos[0] = lp.x, h.light_pos[1] = lp.y, h.light_pos[2] = lp.z;
	w.chunk( "header", &h, sizeof( h ) );
	w.chunk( "blas_table", blasTable.data(), blasTable.size() * sizeof( FlatBlas ) );
	w.chunk( "nodes", nodes.data(), nodes.size() * sizeof( BVHNode ) );
	w.chunk( "tris", tris.data(), tris.size() * sizeof( Tri ) );
	w.chunk( "tri_indices", triIdx.data(), triIdx.size() * sizeof( uint ) );
	w.chunk( "obj_material", objMaterial.data(), objMaterial.size() * sizeof( int32_t ) );
	w.chunk( "materials", mats.data(), mats.size() * sizeof( FlatMaterial ) );
	w.chunk( "tex_table", texTable.data(), texTable.size() * sizeof( FlatTexture ) );
	w.chunk( "tex_pixels", texPixels.data(), texPixels.size() * sizeof( uint ) );
	w.close();
	return 0;
}

// Moves the vertices of one acceleration structure's triangles (v9: 9 floats per triangle, the structure's own triangle
// order) and calls the reference's own Refit() on it (bvh.cpp:26-43 / blas_bvh.cpp:104-121).  BVH builds only.
// A following ref_flatten dumps the refitted nodes: that is what pins oracle/rt_oracle.c orc_refit_bvh.
int ref_refit( int blas, int n, const float* v9 )
{
#if defined(REF_SCENE_FILE) && !defined(USE_KDTree) && !defined(USE_Grid)
	auto& acc = g_renderer->scene.acc;
	(void)blas;
#elif !defined(REF_SCENE_FILE) && !defined(TLAS_USE_KDTree) && !defined(TLAS_USE_Grid)
	if (blas < 0 || blas >= (int)g_renderer->scene.tlas.blas.size()) return -1;
	auto& acc = *g_renderer->scene.tlas.blas[blas];
#endif
#if (defined(REF_SCENE_FILE) && !defined(USE_KDTree) && !defined(USE_Grid)) || (!defined(REF_SCENE_FILE) && !defined(TLAS_USE_KDTree) && !defined(TLAS_USE_Grid))
	if (n != (int)acc.triangles.size()) return -1;
	for (int i = 0; i < n; i++)
	{
		const float* v = v9 + 9 * (size_t)i;
		acc.triangles[i].vertex0 = float3( v[0], v[1], v[2] );
		acc.triangles[i].vertex1 = float3( v[3], v[4], v[5] );
		acc.triangles[i].vertex2 = float3( v[6], v[7], v[8] );
	}
	acc.Refit();
	return 0;
#else
	(void)blas, (void)n, (void)v9;
	return -2;
#endif
}

int ref_sizeof_tri() { return (int)sizeof( Tri ); }
int ref_sizeof_node() { return (int)sizeof( BVHNode ); }

} // extern "C"
