"""Parity of the CUDA path against the oracle, through the C-ABI (run on the B200 box: pytest -m gpu).

Bars (BASELINE.json north_star):
  * hits (t, u, v, objIdx, triIdx) and the traversed/tested work counters: BIT-EXACT for every ray,
    including rays with zero direction components (NaN slab semantics) — no documented ties needed;
  * IsOccluded: equal booleans;
  * Whitted radiance: max abs error <= WHITTED_TOL (top-down weights reassociate a few products);
  * path tracer at equal spp with the reference's RNG: identical ray counts (same path decisions);
    radiance: the accumulator is BIT-IDENTICAL to the reference's after any number of frames, however they were
    launched (one Tick per call, all frames in one call, look-ahead, either schedule, every accelerator): expf /
    atan2f / acosf are glibc's routines restated on the device (tests/test_glibc_math.py) and the samples of a
    multi-frame launch are added to the accumulator in the reference's frame order (k_sum_frames).  Only the
    CUDA-libm build variant (RT_B200_LIB=...cudamath.so, never the default) is held to a tolerance instead:
    per-pixel |diff| <= PT_TOL except PT_FLIP_FRACTION of the pixels (a sky lookup on a texel border can pick the
    neighbouring texel), RMSE / PSNR stated in check_pt.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, all_scene_names, biteq, random_rays, shadow_rays_from

from cpu_ray_tracer_b200 import abi

pytestmark = pytest.mark.gpu

WHITTED_TOL = 2e-5        # absolute, radiance values reach 24
PT_TOL = 1e-4             # absolute per channel, accumulated over the frames rendered
PT_FLIP_FRACTION = 2e-4   # pixels allowed to differ by more (texel flips)
PT_RMSE = 2e-3
HIT_FIELDS = ("t", "u", "v", "obj_idx", "tri_idx", "traversed", "tested")

SCENES = all_scene_names()


def assert_hits_equal(ours, theirs, what):
    for f in HIT_FIELDS:
        a, b = ours[f], theirs[f]
        assert biteq(a, b), f"{what}: {f} differs on {(a.view(np.uint32) != b.view(np.uint32)).sum()} of {len(a)} rays"


@pytest.mark.parametrize("name", SCENES)
def test_find_nearest_primary_bit_exact(name, oracles, gpu_scenes):
    po, sc = oracles(name), gpu_scenes(name)
    W, H = 320, 192
    for cam in (po.camera_default(W, H), po.camera_look_at((1.6, 0.9, -1.4), (0.0, -0.4, 1.0), W, H)):
        rays = po.primary_rays(cam, W, H)
        ref, _ = po.find_nearest(rays)
        got = sc.FindNearest(rays)
        assert (ref["obj_idx"] >= 2).any()
        assert_hits_equal(got, ref, name)


@pytest.mark.parametrize("name", SCENES)
def test_find_nearest_incoherent_and_degenerate_rays(name, oracles, gpu_scenes, flat_scenes):
    """random rays, axis-aligned rays (rD = inf) and origins on node-box planes (0 * inf = NaN)"""
    po, sc = oracles(name), gpu_scenes(name)
    rays = random_rays(flat_scenes(name), 40000, seed=7)
    ref, _ = po.find_nearest(rays)
    assert_hits_equal(sc.FindNearest(rays), ref, name)
    # bounded rays: Ray::t on entry limits every primitive and the BVH cull
    rays["tmax"] = np.random.default_rng(3).uniform(0.05, 6.0, len(rays)).astype(np.float32)
    ref, _ = po.find_nearest(rays)
    assert_hits_equal(sc.FindNearest(rays), ref, name + " (bounded)")


@pytest.mark.parametrize("name", SCENES)
def test_secondary_rays_and_occlusion(name, oracles, gpu_scenes, flat_scenes):
    po, sc = oracles(name), gpu_scenes(name)
    W, H = 256, 160
    rays = po.primary_rays(po.camera_default(W, H), W, H)
    hits, _ = po.find_nearest(rays)
    sr = shadow_rays_from(flat_scenes(name), rays, hits)
    ref, _ = po.is_occluded(sr)
    got = sc.IsOccluded(sr)
    assert np.array_equal(got, ref)
    # the same rays as closest-hit queries (incoherent secondary rays leaving surfaces)
    sr2 = sr.copy()
    sr2["tmax"] = 1e34
    ref2, _ = po.find_nearest(sr2)
    assert_hits_equal(sc.FindNearest(sr2), ref2, name + " (secondary)")


def test_edge_cases_empty_ragged_single(oracles, gpu_scenes, flat_scenes):
    name = "golden_tlas"
    po, sc = oracles(name), gpu_scenes(name)
    assert len(sc.FindNearest(np.zeros(0, abi.RAY_DTYPE))) == 0
    assert len(sc.IsOccluded(np.zeros(0, abi.RAY_DTYPE))) == 0
    rays = random_rays(flat_scenes(name), 4099, seed=11)  # prime length: ragged last warp / CTA
    for n in (1, 31, 33, 127, 129, 4099):
        ref, _ = po.find_nearest(rays[:n])
        assert_hits_equal(sc.FindNearest(rays[:n]), ref, f"n={n}")
        occ, _ = po.is_occluded(rays[:n])
        assert np.array_equal(sc.IsOccluded(rays[:n]), occ)
    # rays that can hit nothing: straight up from above everything
    up = rays[:64].copy()
    up["O"][:, 1] = 50.0
    up["D"] = np.array([0, 1, 0], np.float32)
    got = sc.FindNearest(up)
    assert (got["obj_idx"] == -1).all() and (got["tri_idx"] == -1).all() and biteq(got["t"], up["tmax"])


def test_error_behaviour(flat_scenes):
    import ctypes as C
    from cpu_ray_tracer_b200 import api
    L = api.lib()
    h = C.c_void_p()
    assert L.rt_scene_create(None, 0, 0, C.byref(h)) == abi.RT_ERR_INVALID
    d = flat_scenes("golden_file").desc()
    d.kind = 7
    assert L.rt_scene_create(C.byref(d), 0, 0, C.byref(h)) == abi.RT_ERR_INVALID
    d = flat_scenes("golden_file").desc()
    assert L.rt_scene_create(C.byref(d), 99, 0, C.byref(h)) == abi.RT_ERR_NO_DEVICE
    assert b"no such CUDA device" in L.rt_last_error()
    d = flat_scenes("golden_tlas").desc()
    d.kind = abi.RT_SCENE_FLAT  # a flat scene with several BVHs is rejected
    assert L.rt_scene_create(C.byref(d), 0, 0, C.byref(h)) == abi.RT_ERR_INVALID
    assert L.rt_find_nearest(None, None, None, 4) == abi.RT_ERR_INVALID


def test_error_behaviour_kdtree_and_grid_inputs(flat_scenes):
    """malformed flattened KD-trees / grids are rejected with RT_ERR_INVALID and a message, never traversed"""
    import ctypes as C
    from cpu_ray_tracer_b200 import api
    L = api.lib()
    h = C.c_void_p()

    def create(fs):
        d = fs.desc()
        return L.rt_scene_create(C.byref(d), 0, 0, C.byref(h)), L.rt_last_error()

    kd = flat_scenes("golden_kd").copy()
    interior = int(np.argmax(kd.kd_nodes["left"] >= 0))
    kd.kd_nodes["left"][interior] = len(kd.kd_nodes) + 5
    st, msg = create(kd)
    assert st == abi.RT_ERR_INVALID and b"child index" in msg
    kd = flat_scenes("golden_kd").copy()
    kd.kd_tri_indices[0] = len(kd.tris) + 1
    leaf = int(np.argmax((kd.kd_nodes["left"] < 0) & (kd.kd_nodes["tri_count"] > 0) & (kd.kd_nodes["tri_start"] == 0)))
    assert kd.kd_nodes["tri_start"][leaf] == 0
    st, msg = create(kd)
    assert st == abi.RT_ERR_INVALID and b"triangle index" in msg
    kd = flat_scenes("golden_kd").copy()
    kd.kd_nodes["right"][interior] = 0   # a cycle through the root
    st, msg = create(kd)
    assert st == abi.RT_ERR_INVALID and (b"not a tree" in msg or b"deeper" in msg)
    gr = flat_scenes("golden_grid").copy()
    gr.grid_cell_start[3] = len(gr.grid_tri_indices) + 7
    st, msg = create(gr)
    assert st == abi.RT_ERR_INVALID and b"cell range" in msg
    gr = flat_scenes("golden_grid").copy()
    gr.grid_header["resolution"][0][1] = 0
    st, msg = create(gr)
    assert st == abi.RT_ERR_INVALID and b"resolution" in msg
    d = flat_scenes("golden_tlas_kd").desc()
    d.blas_accel = None
    assert L.rt_scene_create(C.byref(d), 0, 0, C.byref(h)) == abi.RT_ERR_INVALID
    d = flat_scenes("golden_kd").desc()
    d.kd_nodes = None
    assert L.rt_scene_create(C.byref(d), 0, 0, C.byref(h)) == abi.RT_ERR_INVALID


@pytest.mark.parametrize("name", ["golden_file", "golden_tlas", "golden_kd"])
def test_device_pointer_entry_points(name, oracles, gpu_scenes, flat_scenes):
    """rt_find_nearest_device(_ex) / rt_is_occluded_device on buffers owned by torch; the coherence hint changes the kernel, not the hits"""
    import torch
    po, sc = oracles(name), gpu_scenes(name)
    rays = random_rays(flat_scenes(name), 10000, seed=5)
    ref, _ = po.find_nearest(rays)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32)).cuda()
    d_hits = torch.empty((len(rays), 32), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    sc.FindNearestDevice(d_rays.data_ptr(), d_hits.data_ptr(), len(rays), stream)
    got = d_hits.cpu().numpy().reshape(-1).view(abi.HIT_DTYPE)
    assert_hits_equal(got, ref, "device pointers")
    d_hits.zero_()
    sc.FindNearestDevice(d_rays.data_ptr(), d_hits.data_ptr(), len(rays), stream, incoherent=True)   # RT_RAYS_INCOHERENT
    assert_hits_equal(d_hits.cpu().numpy().reshape(-1).view(abi.HIT_DTYPE), ref, "device pointers, incoherent hint")
    d_occ = torch.empty(len(rays), dtype=torch.uint8, device="cuda")
    sc.IsOccludedDevice(d_rays.data_ptr(), d_occ.data_ptr(), len(rays), stream)
    occ, _ = po.is_occluded(rays)
    assert np.array_equal(d_occ.cpu().numpy(), occ)


# ------------------------------------------------------------------------------------------------
# integrators
# ------------------------------------------------------------------------------------------------
def psnr(a, b, peak):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return np.inf if mse == 0 else 10 * np.log10(peak * peak / mse)


def exact_build():
    from cpu_ray_tracer_b200 import api
    return os.path.basename(api.LIB_PATH) == "librt_b200.so"


def check_pt(gacc, oacc, frames, what):
    if exact_build():
        a, b = np.ascontiguousarray(gacc[..., :3]), np.ascontiguousarray(oacc[..., :3], dtype=np.float32)
        bad = (a.view(np.uint32) != b.view(np.uint32)).any(-1)
        assert not bad.any(), f"{what}: accumulator differs from the reference's in {int(bad.sum())} of {bad.size} pixels (max abs {np.nan_to_num(np.abs(a - b)).max()})"
        return
    d = np.abs(gacc[..., :3].astype(np.float64) - oacc[..., :3])
    assert not np.isnan(gacc).any() or np.isnan(oacc).any()
    d = np.nan_to_num(d)
    worst = d.max(-1)
    flips = int((worst > PT_TOL * frames).sum())
    assert flips <= max(2, PT_FLIP_FRACTION * worst.size * frames), f"{what}: {flips} pixels differ by more than {PT_TOL * frames}"
    rmse = float(np.sqrt((d ** 2).mean()))
    assert rmse < PT_RMSE, f"{what}: RMSE {rmse}"
    assert psnr(np.nan_to_num(gacc[..., :3]), np.nan_to_num(oacc[..., :3]), 24.0 * frames) > 70.0


@pytest.mark.parametrize("name", SCENES)
def test_whitted_vs_oracle(name, oracles, gpu_scenes):
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    po, sc = oracles(name), gpu_scenes(name, counters=False)
    W, H = 320, 192
    cam = po.camera_default(W, H)
    oacc, ost = po.render_whitted(cam, porthost.default_params(abi.RT_INTEGRATOR_WHITTED, W, H))
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_WHITTED, W, H).Init()
    r.Tick(0)
    gacc = r.accumulator
    c = r.counters()
    assert c["extension_rays"] == ost["extension_rays"] and c["shadow_rays"] == ost["shadow_rays"]
    d = np.nan_to_num(np.abs(gacc - oacc))
    assert d.max() <= WHITTED_TOL, f"{name}: max abs error {d.max()}"
    same = (gacc.view(np.uint32) == oacc.view(np.uint32)).mean()
    assert same > 0.9, f"{name}: only {same:.3f} of the floats are bit-identical"
    # Tick overwrites (renderer.cpp:155): a second frame gives the same image, not twice the image
    r.Tick(0)
    assert np.nan_to_num(np.abs(r.accumulator - oacc)).max() <= WHITTED_TOL
    # screen->pixels through RGBF32_to_RGB8
    px, opx = r.screen_pixels(), po.to_rgb8(oacc, 1.0)
    chan = lambda p: np.stack([(p >> 16) & 255, (p >> 8) & 255, p & 255], -1).astype(np.int32)
    assert np.abs(chan(px) - chan(opx)).max() <= 1
    r.close()


@pytest.mark.parametrize("schedule", [abi.RT_SCHEDULE_STREAMS, abi.RT_SCHEDULE_WAVEFRONT], ids=["streams", "wavefront"])
@pytest.mark.parametrize("name", SCENES)
def test_path_tracer_vs_oracle_reference_rng(name, schedule, oracles, gpu_scenes):
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    po, sc = oracles(name), gpu_scenes(name, counters=False)
    W, H, frames = 320, 192, 3
    cam = po.camera_default(W, H)
    oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1, frames, 1)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, schedule=schedule).Init()
    for _ in range(frames):
        r.Tick(0)  # one Renderer::Tick per call, spp advances 1, 2, 3
    assert r.spp == 1 + frames
    gacc = r.accumulator
    c = r.counters()
    # same RNG streams + bit-exact hits => every path takes the same decisions => same number of rays
    assert c["extension_rays"] == ost["extension_rays"], (c, ost)
    assert c["paths"] == ost["paths"] == W * H * frames
    check_pt(gacc, oacc, frames, name)
    # the same frames in ONE call (all (tile, frame) streams in flight together)
    r.ClearAccumulator()
    r.reset_counters()
    r.render(frames, first_spp=1)
    assert r.counters()["extension_rays"] == ost["extension_rays"]
    check_pt(r.accumulator, oacc, frames, name + " (batched)")
    r.close()


@pytest.mark.parametrize("schedule", [abi.RT_SCHEDULE_STREAMS, abi.RT_SCHEDULE_WAVEFRONT], ids=["streams", "wavefront"])
@pytest.mark.parametrize("name", ["golden_file", "golden_tlas", "golden_kd", "golden_grid", "golden_tlas_kd", "golden_tlas_grid"])
def test_path_tracer_passes(name, schedule, oracles, gpu_scenes):
    """Renderer::passes > 1: consecutive samples per pixel from the tile's stream, spp advancing by `passes` per Tick"""
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    po, sc = oracles(name), gpu_scenes(name, counters=False)
    W, H, frames, passes = 192, 112, 2, 3
    cam = po.camera_default(W, H)
    oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H, passes=passes), 1, frames, passes)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, schedule=schedule, lookahead_frames=8).Init()
    r.passes = passes
    for _ in range(frames):
        r.Tick(0)
    assert r.spp == 1 + frames * passes
    c = r.counters()
    assert c["extension_rays"] == ost["extension_rays"] and c["paths"] == ost["paths"] == W * H * frames * passes
    check_pt(r.accumulator, oacc, frames * passes, name)
    # the slider moves back to 1 between frames: the next Tick continues at the current spp with one sample per pixel
    r.passes = 1
    r.Tick(0)
    oacc, _ = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1 + frames * passes, 1, 1, accumulator=oacc)
    check_pt(r.accumulator, oacc, frames * passes + 1, name + " (passes back to 1)")
    r.close()


@pytest.mark.parametrize("kind", ["file", "tlas", "kd", "grid", "tlas_kd", "tlas_grid"])
def test_integrators_vs_committed_golden(kind, oracles, gpu_scenes):
    """the vectors the reference's own build produced (tests/golden/make_golden.py), second camera too"""
    from cpu_ray_tracer_b200 import api
    g = np.load(os.path.join(GOLDEN, f"golden_{kind}.npz"))
    W, H, frames = (int(x) for x in g["meta"])
    sc = gpu_scenes(f"golden_{kind}", counters=False)
    for c in (0, 1):
        wh = api.GpuRenderer(sc, abi.RT_INTEGRATOR_WHITTED, W, H).Init()
        pt = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
        if c == 1:
            wh.camera.SetCameraState(g["look_at"][0], g["look_at"][1])
            pt.camera.SetCameraState(g["look_at"][0], g["look_at"][1])
        wh.Tick(0)
        assert np.nan_to_num(np.abs(wh.accumulator - g[f"whitted{c}"])).max() <= WHITTED_TOL
        pt.render(frames)
        check_pt(pt.accumulator, g[f"pt{c}"], frames, f"golden {kind} cam{c}")
        wh.close(), pt.close()


@pytest.mark.parametrize("schedule", [abi.RT_SCHEDULE_STREAMS, abi.RT_SCHEDULE_WAVEFRONT], ids=["streams", "wavefront"])
@pytest.mark.parametrize("name,passes", [("golden_file", 1), ("golden_tlas", 2), ("golden_kd", 1)])
def test_path_tracer_per_pixel_seed_mode(name, passes, schedule, oracles, gpu_scenes):
    """RT_SEED_PER_PIXEL: one stream per pixel per frame (shared by the pixel's `passes` samples); the stream kernel runs
    every pixel as its own stream, the wavefront as before"""
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    po, sc = oracles(name), gpu_scenes(name, counters=False)
    W, H, frames = 128, 80, 2
    cam = po.camera_default(W, H)
    oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H, seed_mode=abi.RT_SEED_PER_PIXEL, passes=passes), 1, frames, passes)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, seed_mode=abi.RT_SEED_PER_PIXEL, schedule=schedule).Init()
    r.passes = passes
    r.render(frames)
    c = r.counters()
    assert c["extension_rays"] == ost["extension_rays"] and c["paths"] == ost["paths"] == W * H * frames * passes
    check_pt(r.accumulator, oacc, frames * passes, "per-pixel seeds")
    r.close()
    if passes == 1:
        # one Tick per call with look-ahead frames and interleaved tile shards: same image
        parts = [api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, seed_mode=abi.RT_SEED_PER_PIXEL, schedule=schedule,
                                 lookahead_frames=4, tile_begin=k, tile_end=(W // 16) * (H // 16), tile_step=2).Init() for k in range(2)]
        for q in parts:
            for _ in range(frames):
                q.Tick(0)
        check_pt(parts[0].accumulator + parts[1].accumulator, oacc, frames, "per-pixel seeds, look-ahead, interleaved tiles")
        for q in parts:
            q.close()


def test_depth_limit_and_partial_tiles(oracles, gpu_scenes):
    """depthLimit other than 5; a height that is not a multiple of 16 renders only whole tiles"""
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    name = "golden_tlas"
    po, sc = oracles(name), gpu_scenes(name, counters=False)
    W, H = 144, 90  # 9 x 5 tiles, bottom 10 rows never rendered by the path tracer
    cam = po.camera_default(W, H)
    for depth in (0, 1, 3):
        oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H, depth_limit=depth), 1, 2, 1)
        r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, depthLimit=depth).Init()
        r.render(2)
        assert r.counters()["extension_rays"] == ost["extension_rays"]
        gacc = r.accumulator
        check_pt(gacc, oacc, 2, f"depth {depth}")
        assert (gacc[80:] == 0).all() and (oacc[80:] == 0).all()
        r.close()
        ow, _ = po.render_whitted(cam, porthost.default_params(abi.RT_INTEGRATOR_WHITTED, W, H, depth_limit=depth))
        r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_WHITTED, W, H, depthLimit=depth).Init()
        r.Tick(0)
        assert np.nan_to_num(np.abs(r.accumulator - ow)).max() <= WHITTED_TOL
        r.close()


def test_sharding_properties_full_size(gpu_scenes, oracles):
    """BASELINE size (1920x1080): size-independent properties the multi-GPU split relies on.
       frames:  render(spp 1..4) == render(1,3) + render(2,4)      (sample-index sharding, stride 2)
       tiles:   render(all tiles) == render(first half) + render(second half)   (tile sharding)
       rays:    the ray count is a deterministic function of (scene, camera, spp range)
    Tile shards and the two schedules give the image bit for bit (disjoint pixels, frame-ordered sums); sample-index shards add
    the same samples in another order: float reassociation only."""
    from cpu_ray_tracer_b200 import api
    from conftest import baked_scenes
    name = "wok_teapot_flat" if "wok_teapot_flat" in baked_scenes() else "golden_file"
    sc = gpu_scenes(name, counters=False)
    W, H = 1920, 1080
    full = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
    full.render(4, first_spp=1)
    ref = full.accumulator
    rays_full = full.counters()["extension_rays"]
    # the two schedules trace the same rays and build the same image
    wf = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, schedule=abi.RT_SCHEDULE_WAVEFRONT).Init()
    wf.render(4, first_spp=1)
    assert wf.counters()["extension_rays"] == rays_full
    assert biteq(wf.accumulator, ref)  # frame-ordered accumulation in both schedules
    wf.close()
    assert (ref[1072:] == 0).all()  # 1080 % 16 = 8 rows never rendered
    assert ref[:1072, :, :3].sum() > 0
    # sample-index sharding
    a = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
    a.render(2, first_spp=1, stride=2)
    a.render(2, first_spp=2, stride=2)
    assert a.counters()["extension_rays"] == rays_full
    assert np.abs(a.accumulator - ref).max() <= 1e-4
    a.close()
    # tile sharding
    tiles = (W // 16) * (H // 16)
    lo = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, tile_begin=0, tile_end=tiles // 2).Init()
    hi = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, tile_begin=tiles // 2, tile_end=tiles).Init()
    lo.render(4, first_spp=1), hi.render(4, first_spp=1)
    assert lo.counters()["extension_rays"] + hi.counters()["extension_rays"] == rays_full
    la, ha = lo.accumulator, hi.accumulator
    assert not ((la[..., :3].sum(-1) != 0) & (ha[..., :3].sum(-1) != 0)).any()  # disjoint pixels
    assert biteq(la + ha, ref)
    lo.close(), hi.close()
    # interleaved tile sharding (tile_step): three "ranks", both schedules
    for sched in (abi.RT_SCHEDULE_STREAMS, abi.RT_SCHEDULE_WAVEFRONT):
        parts = [api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, tile_begin=k, tile_end=tiles, tile_step=3, schedule=sched).Init() for k in range(3)]
        for q in parts:
            q.render(4, first_spp=1)
        assert sum(q.counters()["extension_rays"] for q in parts) == rays_full
        accs = [q.accumulator for q in parts]
        hit = sum((x[..., :3].sum(-1) != 0).astype(np.int32) for x in accs)
        assert hit.max() <= 1                                                         # disjoint pixels
        assert biteq(accs[0] + accs[1] + accs[2], ref)
        for q in parts:
            q.close()
    # oracle spot check at full width on the top tile rows
    po = oracles(name)
    from oracle import porthost
    p = porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H)
    p.tile_begin, p.tile_end = 0, 2 * (W // 16)
    oacc, _ = po.render_pt(po.camera_default(W, H), p, 1, 4, 1)
    check_pt(ref[:32], oacc[:32], 4, "1080p top rows")
    full.close()


@pytest.mark.parametrize("mode", ["simple", "voted"])
def test_alternative_traversal_kernels_agree(mode, monkeypatch, oracles, flat_scenes):
    """RT_B200_TRAVERSAL=simple selects the one-thread-per-ray kernels, =voted the ray-queue traversal with the stream
    kernel's action vote (both kept for A/B profiling); every kernel family must give the oracle's bits"""
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    monkeypatch.setenv("RT_B200_TRAVERSAL", mode)
    for name in ("golden_file", "golden_tlas") + (("golden_kd", "golden_tlas_grid") if mode == "simple" else ()):
        po = oracles(name)
        flat = flat_scenes(name)
        sc = api.open_scene(flat, counters=True)
        rays = random_rays(flat, 20000, seed=21)
        ref, _ = po.find_nearest(rays)
        assert_hits_equal(sc.FindNearest(rays), ref, name + f" ({mode} kernels)")
        occ, _ = po.is_occluded(rays)
        assert np.array_equal(sc.IsOccluded(rays), occ)
        W, H = 128, 80
        cam = po.camera_default(W, H)
        oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1, 2, 1)
        r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
        r.render(2)
        assert r.counters()["extension_rays"] == ost["extension_rays"]
        check_pt(r.accumulator, oacc, 2, name + " (simple kernels)")
        r.close()
        ow, _ = po.render_whitted(cam, porthost.default_params(abi.RT_INTEGRATOR_WHITTED, W, H))
        r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_WHITTED, W, H).Init()
        r.Tick(0)
        assert np.nan_to_num(np.abs(r.accumulator - ow)).max() <= WHITTED_TOL
        r.close()
        sc.close()


def test_depth_limit_above_stream_kernel_capacity_uses_wavefront(oracles, gpu_scenes):
    """depthLimit > 8 does not fit the stream kernel's register-resident throughput stack: the renderer falls
    back to the wavefront schedule and still matches the oracle"""
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    name = "golden_file"
    po, sc = oracles(name), gpu_scenes(name, counters=False)
    W, H, frames, depth = 96, 64, 2, 11
    cam = po.camera_default(W, H)
    oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H, depth_limit=depth), 1, frames, 1)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, depthLimit=depth).Init()
    r.render(frames)
    c = r.counters()
    assert c["extension_rays"] == ost["extension_rays"] and c["wavefront_iterations"] > 0
    check_pt(r.accumulator, oacc, frames, "depth 11")
    r.close()


def test_large_batches_counters_and_profiling_api(oracles, gpu_scenes, flat_scenes):
    """a batch larger than one launch chunk, counter reset, per-stage times, path-tracer display scale"""
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    name = "golden_tlas"
    po, sc = oracles(name), gpu_scenes(name)
    base = random_rays(flat_scenes(name), 1 << 16, seed=3)
    rays = np.tile(base, 40)                      # 2.6 M rays through one rt_find_nearest call
    got = sc.FindNearest(rays)
    ref, _ = po.find_nearest(base)
    for k in (0, 17, 39):
        assert_hits_equal(got[k * len(base):(k + 1) * len(base)], ref, f"copy {k}")
    occ = sc.IsOccluded(rays)
    oref, _ = po.is_occluded(base)
    assert np.array_equal(occ.reshape(40, -1), np.broadcast_to(oref, (40, len(base))))
    W, H = 128, 80
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, schedule=abi.RT_SCHEDULE_WAVEFRONT).Init()
    r.set_profiling(True)
    r.render(2)
    st = r.stage_times()
    assert st["extend"][1] > 0 and st["shade"][1] == st["extend"][1] and st["extend"][0] > 0
    r.set_profiling(False)
    c = r.counters()
    assert c["paths"] == W * H * 2 and c["kernel_launches"] > 0
    r.reset_counters()
    assert r.counters()["extension_rays"] == 0 and r.counters()["paths"] == 0
    # screen->pixels: accumulator / (spp + passes) as the reference displays it after 2 Ticks from spp = 1
    acc = r.accumulator
    px = r.screen_pixels(scale=1.0 / 3.0)
    opx = po.to_rgb8(acc, 1.0 / 3.0)
    assert np.array_equal(px, opx)
    r.close()


def test_renderer_argument_errors(gpu_scenes):
    import ctypes as C
    from cpu_ray_tracer_b200 import api
    L = api.lib()
    sc = gpu_scenes("golden_file")
    p = abi.rt_render_params()
    L.rt_render_params_default(C.byref(p), abi.RT_INTEGRATOR_PATH, 64, 64)
    assert p.depth_limit == 5 and abs(p.epsilon - 0.001) < 1e-9 and p.seed_mode == abi.RT_SEED_REFERENCE_TILE
    h = C.c_void_p()
    p.width = 0
    assert L.rt_renderer_create(sc.handle, C.byref(p), C.byref(h)) == abi.RT_ERR_INVALID
    p.width, p.integrator = 64, 9
    assert L.rt_renderer_create(sc.handle, C.byref(p), C.byref(h)) == abi.RT_ERR_INVALID
    assert L.rt_renderer_create(None, C.byref(p), C.byref(h)) == abi.RT_ERR_INVALID
    assert L.rt_renderer_render(None, 1, 1, 1) == abi.RT_ERR_INVALID
    out = C.c_double()
    assert L.rt_measure_gather_bandwidth(0, 16, 1, C.byref(out)) == abi.RT_ERR_INVALID
    assert L.rt_measure_gather_bandwidth(99, 1 << 20, 1, C.byref(out)) == abi.RT_ERR_NO_DEVICE
    assert api.measure_gather_bandwidth(4 << 20) > 100.0


def test_lookahead_ticks_reproduce_the_tick_sequence(oracles, gpu_scenes):
    """lookahead_frames: frames are rendered ahead in one launch and revealed one per Tick, summed in spp order —
    after EVERY Tick the accumulator is what the reference holds after that many Ticks; a camera change discards
    the frames rendered ahead"""
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    name = "golden_tlas"
    po, sc = oracles(name), gpu_scenes(name, counters=False)
    W, H, ticks, L = 128, 80, 7, 4
    cam = po.camera_default(W, H)
    p = porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, lookahead_frames=L).Init()
    single = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()   # one frame per launch, no look-ahead
    oacc, gsum = None, np.zeros((H, W, 4), np.float32)
    for k in range(ticks):
        oacc, _ = po.render_pt(cam, p, 1 + k, 1, 1, accumulator=oacc)
        r.Tick(0)
        g = r.accumulator
        check_pt(g, oacc, k + 1, f"tick {k + 1}")
        # the floats that differ from the oracle differ by libm ulps (expf, atan2f, acosf), not by summation order
        same = (g[..., :3].view(np.uint32) == oacc[..., :3].view(np.uint32)).mean()
        assert same > 0.95, f"tick {k + 1}: only {same:.4f} bit-identical"
        # bit-exact against the frames rendered one by one and summed in spp order on the host
        single.ClearAccumulator()
        single.render(1, first_spp=1 + k)
        gsum = gsum + single.accumulator
        assert biteq(g[..., :3], gsum[..., :3]), f"tick {k + 1}: look-ahead accumulator differs from the ordered sum of single frames"
    single.close()
    assert r.spp == 1 + ticks
    # 7 Ticks with 4 frames per launch: two launches rendered 8 frames
    assert r.counters()["paths"] == W * H * 8
    # camera change: the frame rendered ahead for spp 8 with the old camera must not be used
    look = ((1.6, 0.9, -1.4), (0.0, -0.4, 1.0))
    r.camera.SetCameraState(*look)
    r.ClearAccumulator()
    r.Tick(0)
    cam2 = po.camera_look_at(look[0], look[1], W, H)
    o2, _ = po.render_pt(cam2, p, 1 + ticks, 1, 1)
    check_pt(r.accumulator, o2, 1, "after camera change")
    r.close()
