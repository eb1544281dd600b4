"""Frame-level parity: the CUDA wavefront path against the CPU oracle on identical scenes, cameras and RNG seeds.

Bars (BASELINE.json north_star): primary hit (instance, primitive, t) bit-exact; per-pixel radiance within a relative
L1 error of 1e-3 (the device libm differs from glibc in sinf/cosf/logf/expf/powf by a few ulp; a Russian-roulette or
reservoir decision that flips on such a difference changes isolated pixels, which is what the L1 bar absorbs)."""
import numpy as np
import pytest

import lumenrenderer_b200 as lr
from lumenrenderer_b200 import api, scenes
from conftest import rel_l1

pytestmark = pytest.mark.gpu

RADIANCE_TOL = 1e-3


def _pair(oracle, scene, **kw):
    st = lr.Settings(**kw)
    g = lr.Renderer(st); c = api.Renderer(oracle, st)
    g.load_scene(scene); c.load_scene(scene)
    return g, c


def _check_hits(g, c):
    hg, hc = g.read_primary_hits(), c.read_primary_hits()
    assert np.array_equal(hg["t"] > 0, hc["t"] > 0)
    for f in ("instance", "primitive", "t", "u", "v"):
        assert np.array_equal(hg[f], hc[f]), f"primary hit field {f} differs in {(hg[f] != hc[f]).sum()} pixels"


def test_cornell_c1_nee(oracle):
    """BASELINE config C1: Cornell box 256x256, 1 spp, 1 bounce (depth 2), no ReSTIR."""
    g, c = _pair(oracle, scenes.cornell_box(), width=256, height=256, depth=2, restir=False)
    g.render_frames(1); c.render_frames(1)
    _check_hits(g, c)
    sg, sc = g.read_surface(), c.read_surface()
    assert np.array_equal(sg[..., :8], sc[..., :8]), "primary surface position/t/normal/flags differ"
    assert np.array_equal(sg, sc), "primary surface records differ"
    assert np.array_equal(g.read_motion_vectors(), c.read_motion_vectors())
    for ch in range(4):
        assert rel_l1(g.read_channel(ch)[..., :3], c.read_channel(ch)[..., :3]) < RADIANCE_TOL or np.abs(c.read_channel(ch)).sum() == 0
    assert rel_l1(g.read_hdr()[..., :3], c.read_hdr()[..., :3]) < RADIANCE_TOL
    cg, cc = g.frame_counters(), c.frame_counters()
    assert cg["extend_rays"] == cc["extend_rays"] and cg["shadow_rays"] == cc["shadow_rays"]
    assert np.abs(g.read_ldr().astype(int) - c.read_ldr().astype(int)).max() <= 1
    g.close(); c.close()


@pytest.mark.parametrize("temporal,spatial", [(False, False), (True, False), (True, True)])
def test_cornell_restir(oracle, temporal, spatial):
    g, c = _pair(oracle, scenes.cornell_box(), width=160, height=120, depth=3, restir=True, restir_temporal=temporal, restir_spatial=spatial)
    for frame in range(3):
        g.render_frames(1); c.render_frames(1)
        _check_hits(g, c)
        err = rel_l1(g.read_hdr()[..., :3], c.read_hdr()[..., :3])
        assert err < RADIANCE_TOL, f"frame {frame}: {err}"
    rg, rc = g.read_reservoirs(), c.read_reservoirs()
    assert rel_l1(rg[..., 0], rc[..., 0]) < RADIANCE_TOL         # weight sums
    assert (rg[..., 2] != rc[..., 2]).mean() < 1e-3              # sample counts
    cg, cc = g.frame_counters(), c.frame_counters()
    assert abs(cg["visibility_rays"] - cc["visibility_rays"]) <= 4
    g.close(); c.close()


def test_restir_unbiased_combine(oracle):
    """LbSettings::restir_unbiased = the reference's ReSTIRSettings::enableBiased = false: temporal and spatial reuse take the CombineUnbiased
    branches (ReSTIRKernels.cu:905-970, :1123-1198; pinned against the reference compiled with that constant flipped,
    tests/test_oracle_golden.py). The branch is dead in the shipped reference for a reason: its spatial normalisation counts the stale sample
    count of the output buffer (:951), 0 on a first frame, so weights come out as weightSum / FLT_EPSILON^2 and the history diverges to inf / NaN
    within a few frames — reproduced as written. Compared here: the reservoirs after ONE frame (finite), pixel by pixel."""
    scene = scenes.material_gallery()
    g, c = _pair(oracle, scene, width=160, height=120, depth=2, restir=True, restir_unbiased=True)
    biased = lr.Renderer(lr.Settings(width=160, height=120, depth=2, restir=True)); biased.load_scene(scene)
    for r in (g, c, biased):
        r.render_frames(1)
    _check_hits(g, c)
    assert rel_l1(g.read_hdr()[..., :3], c.read_hdr()[..., :3]) < RADIANCE_TOL          # the frame itself only sees RIS + visibility
    rg, rc, rb = g.read_reservoirs(), c.read_reservoirs(), biased.read_reservoirs()
    assert np.isfinite(rc[..., :4]).all() and np.isfinite(rg[..., :4]).all()
    assert (rg[..., 2] != rc[..., 2]).mean() < 5e-3                                       # sample counts
    close = np.isclose(rg[..., 1], rc[..., 1], rtol=2e-3, atol=1e-6)
    assert close.mean() > 0.99, f"{(~close).sum()} of {close.size} reservoir weights differ"
    assert (rc[..., 1] > 1e6).mean() > 0.01 and not (rb[..., 1] > 1e6).any()              # the unbiased normalisation really ran
    g.close(); c.close(); biased.close()


def test_depth_24_uses_its_own_tickets(oracle):
    """The deepest schedule lb_create accepts: 24 waves of extend + shadow (+ 5 ReSTIR launches) take 53 device tickets, with media in the
    default LB_VOLUME_COMPAT mode 77 — more than the 56 of round 1 (the ticket block now holds 120 and the host refuses a schedule that
    would not fit). Rays of every wave must be the oracle's."""
    g, c = _pair(oracle, scenes.fog_room(), width=96, height=64, depth=24, restir=True)
    for frame in range(2):
        g.render_frames(1); c.render_frames(1)
    _check_hits(g, c)
    cg, cc = g.frame_counters(), c.frame_counters()
    assert cg["extend_rays"] == cc["extend_rays"] and cg["stack_overflows"] == 0
    assert rel_l1(g.read_hdr()[..., :3], c.read_hdr()[..., :3]) < 5e-3
    g.close(); c.close()


@pytest.mark.parametrize("width,height", [(131, 77), (33, 9), (7, 5), (300, 1)])
def test_restir_at_ragged_sizes(oracle, width, height):
    """Pixel counts that are not multiples of the 32-pixel rows, 256-pixel bag groups and 32x8 tiles the ReSTIR kernels hand out
    (partial rows, a single partial group, fewer pixels than one warp, a one-row image)."""
    g, c = _pair(oracle, scenes.material_gallery(), width=width, height=height, depth=3, restir=True)
    for frame in range(3):
        g.render_frames(1); c.render_frames(1)
    _check_hits(g, c)
    assert np.isfinite(g.read_hdr()).all()
    # few pixels: one flipped discrete decision moves the L1 ratio a lot more than at full size, so compare the two estimates pixel-wise
    hg, hc = g.read_hdr()[..., :3], c.read_hdr()[..., :3]
    close = np.isclose(hg, hc, rtol=1e-3, atol=1e-5).all(axis=-1)
    assert close.mean() > 0.99, f"{(~close).sum()} of {close.size} pixels differ"
    rg, rc = g.read_reservoirs(), c.read_reservoirs()
    assert (rg[..., 2] != rc[..., 2]).mean() < 5e-3
    g.close(); c.close()


def test_gallery_all_lobes(oracle):
    """Every Disney lobe, textures (bilinear + sRGB), normal maps, alpha cut-out, override emission/materials, 2 frames of ReSTIR."""
    g, c = _pair(oracle, scenes.material_gallery(), width=192, height=128, depth=4, restir=True)
    for frame in range(2):
        g.render_frames(1); c.render_frames(1)
    _check_hits(g, c)
    sg, sc = g.read_surface(), c.read_surface()
    assert np.array_equal(sg, sc)
    lg, lc = g.read_lights(), c.read_lights()
    assert np.array_equal(lg[0], lc[0]) and np.array_equal(lg[1], lc[1]), "light list / CDF differ"
    err = rel_l1(g.read_hdr()[..., :3], c.read_hdr()[..., :3])
    assert err < 5e-3, err          # a glass + cut-out scene has more discrete decisions per path; see DESIGN.md
    g.close(); c.close()


def test_progressive_blend_and_resolve(oracle):
    g, c = _pair(oracle, scenes.cornell_box(), width=64, height=64, depth=3, restir=False, blend_output=True)
    g.render_frames(4); c.render_frames(4)
    assert rel_l1(g.read_hdr()[..., :3], c.read_hdr()[..., :3]) < RADIANCE_TOL
    ptr, nbytes, frames = g.accum_buffer()
    assert frames == 4 and nbytes == 64 * 64 * 16 and ptr
    before = g.read_hdr().copy()
    g.resolve_accum(4)
    assert np.allclose(g.read_hdr(), before, rtol=1e-6, atol=1e-7)
    g.close(); c.close()


@pytest.mark.parametrize("mode", [lr.VOLUME_COMPAT, lr.VOLUME_DELTA])
@pytest.mark.parametrize("restir", [False, True])
def test_fog_room_volumes(oracle, mode, restir):
    """Config C3 geometry at test size: homogeneous box + heterogeneous 64^3 grid. COMPAT = the reference's 5-step constant-density
    march (GPUVolumetricShadeDirect.cu:8-101); DELTA = delta tracking + ratio-tracked shadow transmittance (identical RNG streams)."""
    g, c = _pair(oracle, scenes.fog_room(grid=32), width=160, height=96, depth=3, restir=restir, volume_mode=mode)
    for frame in range(2):
        g.render_frames(1); c.render_frames(1)
    _check_hits(g, c)
    hg, hc = g.read_hdr()[..., :3], c.read_hdr()[..., :3]
    assert np.isfinite(hg).all()
    assert rel_l1(hg, hc) < 2e-3, rel_l1(hg, hc)
    vg, vc = g.read_channel(lr.CHANNEL_VOLUMETRIC), c.read_channel(lr.CHANNEL_VOLUMETRIC)
    if mode == lr.VOLUME_COMPAT:
        assert np.abs(vc[..., 3]).sum() > 0 and rel_l1(vg[..., 3], vc[..., 3]) < 1e-5       # accumulated density -> alpha
        assert rel_l1(vg[..., :3], vc[..., :3]) < 1e-4
    cg, cc = g.frame_counters(), c.frame_counters()
    assert abs(cg["shadow_rays"] - cc["shadow_rays"]) <= max(4, cc["shadow_rays"] // 2000)
    g.close(); c.close()


def test_overlap_mode_is_bit_identical_to_the_serialised_frame():
    """lb_set_overlap: the ReSTIR chain and the bounce chain of a frame on two streams (default) produce exactly the image, reservoirs
    and ray counts of the serialised frame — they touch disjoint buffers (DIRECT channel vs. the others) — and so does the frame whose late
    bounce waves run as one path-per-lane launch."""
    scene = scenes.material_gallery()
    outs = []
    for overlap in (13, 5, 9, 8, 4, 3, 1, 0):          # 5 is the default; bit 3: the waves from the third on as one launch, a lane per path (k_tail)
        g = lr.Renderer(lr.Settings(width=256, height=160, depth=4, restir=True))
        g.load_scene(scene); g.set_overlap(overlap)
        g.render_frames(3)
        stats = g.frame_stats()
        assert ("restir_join" in stats) == bool(overlap & 6) and all(v >= 0 for v in stats.values())
        outs.append((g.read_hdr(), g.read_reservoirs(), [g.read_channel(c) for c in range(4)], g.frame_counters()))
        g.close()
    b = outs[-1]
    for a in outs[:-1]:
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        for ca, cb in zip(a[2], b[2]):
            assert np.array_equal(ca, cb)
        for k in ("extend_rays", "shadow_rays", "visibility_rays"):
            assert a[3][k] == b[3][k]
    assert outs[0][3]["kernel_launches"] < b[3]["kernel_launches"]          # the fused tail replaces the per-wave launches of waves 2 and 3


@pytest.mark.parametrize("width,height", [(333, 141), (33, 9)])
def test_tma_staged_spatial_pass_is_bit_identical(monkeypatch, width, height):
    """LB_SPATIAL_TMA=1: the spatial-reuse pass with the similarity records of a tile's neighbourhood staged in shared memory by the TMA unit
    (k_spatial_tma; off by default, it measured slower) returns the reservoirs and the image of the gathering kernel bit for bit — on a size
    that is no multiple of the 32 x 16 tile and on one smaller than the 92 x 76 box, so the zero-filled border of the box is exercised too."""
    scene = scenes.material_gallery()
    outs = []
    for tma in ("0", "1"):
        monkeypatch.setenv("LB_SPATIAL_TMA", tma)
        g = lr.Renderer(lr.Settings(width=width, height=height, depth=3, restir=True))
        g.load_scene(scene)
        g.render_frames(3)
        outs.append((g.read_hdr(), g.read_reservoirs()))
        g.close()
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][0], outs[1][0])
    assert np.isfinite(outs[1][0]).all() and outs[1][0].max() > 0


def test_async_readback_equals_blocking_readback():
    """lb_read_hdr_async + lb_readback_wait deliver the same frame as lb_read_hdr, also while the next frame is being rendered."""
    g = lr.Renderer(lr.Settings(width=128, height=96, depth=3, restir=True))
    g.load_scene(scenes.cornell_box())
    bufs = [np.zeros((96, 128, 4), np.float32) for _ in range(2)]
    want = []
    for k in range(4):
        g.render_frames(1)
        g.readback_wait()
        if k:
            assert np.array_equal(bufs[(k - 1) % 2], want[k - 1]), f"frame {k - 1} arrived altered"
        g.read_hdr_async(bufs[k % 2].ctypes.data, bufs[k % 2].nbytes)
        want.append(None)
        # the blocking read of the same frame (the pending copy and this one read the same buffer)
        want[k] = g.read_hdr().copy()
    g.readback_wait()
    assert np.array_equal(bufs[3 % 2], want[3])
    g.close()


def test_gbuffer_side_outputs_bit_exact(oracle):
    """Depth / normal+roughness / albedo side outputs against the oracle's restatement of GPUExtractNRD_DLSSdata.cu."""
    g, c = _pair(oracle, scenes.material_gallery(), width=160, height=96, depth=2, restir=False)
    for r in (g, c):
        r.set_camera_min_max_distance(0.5, 40.0)
        r.render_frames(1)
    (dg, ng, ag), (dc, nc, ac) = g.read_gbuffer(), c.read_gbuffer()
    assert np.array_equal(dg, dc) and dg.max() > 0 and dg.max() <= 1.0
    assert np.array_equal(ng, nc) and np.array_equal(ag, ac)
    assert np.abs(np.linalg.norm(ng[..., :3], axis=-1)[dg > 0] - 1).max() < 2e-3       # fp16-rounded unit normals
    g.close(); c.close()


def test_dynamic_scene_moving_camera_and_instance(oracle):
    """SURVEY 8f-4 (dynamic scenes): the camera dollies and turns and one mesh instance moves between frames. The scene commit
    (flattening, BVH, light list) is redone on the device, motion vectors are non-zero, and temporal reuse reprojects into the
    previous frame's reservoirs (ReSTIRKernels.cu:1044-1048) — all against the oracle, frame by frame."""
    scene = scenes.cornell_box()
    g, c = _pair(oracle, scene, width=160, height=120, depth=3, restir=True)
    cam = scene.camera
    p0 = np.asarray(cam["position"], np.float64)
    moved = len(scene.instances) - 1
    base = np.asarray(scene.instances[moved].get("transform", np.eye(4)), np.float32).reshape(4, 4).copy()
    nonzero_mv = 0
    for frame in range(4):
        ang = 0.03 * frame
        w0, x0, y0, z0 = cam.get("rotation", (1.0, 0.0, 0.0, 0.0)); w1, y1 = float(np.cos(ang / 2)), float(np.sin(ang / 2))   # base * yaw
        rot = (w0 * w1 - y0 * y1, x0 * w1 - z0 * y1, y0 * w1 + w0 * y1, z0 * w1 + x0 * y1)
        m = base.copy(); m[:3, 3] += np.float32(0.02 * frame) * np.array([1, 0.5, -0.25], np.float32)
        for r in (g, c):
            r.set_camera(p0 + np.array([0.02, 0.01, -0.03]) * frame, rot, cam.get("fov_y"))
            if frame:
                r.set_instance_transform(moved, m)
            r.render_frames(1)
        _check_hits(g, c)
        assert np.array_equal(g.read_surface(), c.read_surface())
        mg, mc = g.read_motion_vectors(), c.read_motion_vectors()
        assert np.array_equal(mg, mc)
        nonzero_mv += int((mg != 0).any())
        err = rel_l1(g.read_hdr()[..., :3], c.read_hdr()[..., :3])
        assert err < RADIANCE_TOL, f"frame {frame}: {err}"
    assert nonzero_mv >= 3
    rg, rc = g.read_reservoirs(), c.read_reservoirs()
    assert (rg[..., 2] != rc[..., 2]).mean() < 1e-3
    assert g.frame_counters()["bvh_refits"] == 3          # the three instance moves refitted the hierarchies (no rebuild), and stayed bit-exact
    g.close(); c.close()


def test_split_references_return_the_hits_of_whole_triangles(gpu, monkeypatch):
    """LB_BVH_SPLIT=k (early split clipping: large triangles enter the build as several clipped references, csrc/lb_bvh.cu k_split): hit
    records of 200 K random rays — closest-hit and any-hit — are bit-identical to the hierarchy over whole triangles, on the Cornell box (12
    wall triangles that span the scene: every one is cut into hundreds of references) and on the material gallery; a refit of the split
    hierarchy (a reference then gets its whole triangle's box) keeps them."""
    rng = np.random.default_rng(4)
    o = rng.uniform(-1, 1, (200_000, 3)).astype(np.float32)
    d = rng.normal(size=(200_000, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:1000] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 1000)] * rng.choice([-1.0, 1.0], (1000, 1)).astype(np.float32)      # along the cutting planes
    tmax = (rng.random(200_000) * 4).astype(np.float32)
    st = lr.Settings(width=64, height=48, depth=2, restir=True)
    for scene_fn in (scenes.cornell_box, scenes.material_gallery):
        out = []
        for split in ("0", "16", "64"):
            monkeypatch.setenv("LB_BVH_SPLIT", split)
            with api.Renderer(gpu, st) as r:
                sc = scene_fn()
                r.load_scene(sc); r.render_frames(1)
                fc = r.frame_counters(); assert fc["stack_overflows"] == 0
                out.append((r.trace_closest(o, d), r.trace_any(o, d, tmax), r.read_hdr(), fc["bvh_bytes"]))
                if split == "64":
                    t0 = sc.instances[0].get("transform")            # the same transform again: a transform-only change, nothing moves
                    r.set_instance_transform(0, np.eye(4, dtype=np.float32) if t0 is None else np.asarray(t0, np.float32).reshape(4, 4)); r.render_frames(1)
                    assert r.frame_counters()["bvh_refits"] == 1
                    out.append((r.trace_closest(o, d), r.trace_any(o, d, tmax), None, 0))
        assert out[1][3] > out[0][3]                                 # references were added
        for other in out[1:]:
            assert np.array_equal(out[0][0], other[0]) and np.array_equal(out[0][1], other[1])
        assert np.array_equal(out[0][2], out[1][2]) and np.array_equal(out[0][2], out[2][2])
        assert (out[0][0]["t"] > 0).mean() > 0.3


def test_refitted_hierarchy_returns_the_hits_of_a_rebuilt_one(gpu):
    """bvh_refit (instances moved -> same topology, new boxes) against a full rebuild of the same final scene: hit records of 200 K random
    rays bit-identical, closest-hit and any-hit; instances are moved far (boxes that were disjoint at build time now overlap) and back."""
    scene = scenes.material_gallery()
    rng = np.random.default_rng(9)
    o = rng.uniform(-3, 3, (200_000, 3)).astype(np.float32); o[:, 1] = np.abs(o[:, 1])
    d = rng.normal(size=(200_000, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    tmax = (rng.random(200_000) * 8).astype(np.float32)
    st = lr.Settings(width=64, height=48, depth=2, restir=True)
    moves = [(k, scenes.translate(*(rng.uniform(-1.5, 1.5, 3) * [1, 0.2, 1]), 1.0, float(rng.uniform(0, 360)))) for k in range(1, min(6, len(scene.instances)))]
    with api.Renderer(gpu, st) as a:
        a.load_scene(scene); a.render_frames(1)
        for step in range(3):
            for k, m in moves:
                mm = np.asarray(m, np.float32).reshape(4, 4).copy(); mm[:3, 3] *= np.float32(step + 1)
                a.set_instance_transform(k, mm)
            a.render_frames(1)
        ca = a.frame_counters()
        assert ca["bvh_refits"] == 3 and ca["stack_overflows"] == 0
        ha, oa, la = a.trace_closest(o, d), a.trace_any(o, d, tmax), a.read_lights()
    moved_scene = scenes.material_gallery()
    for k, m in moves:
        mm = np.asarray(m, np.float32).reshape(4, 4).copy(); mm[:3, 3] *= np.float32(3)
        moved_scene.instances[k]["transform"] = mm
    with api.Renderer(gpu, st) as b:
        b.load_scene(moved_scene); b.render_frames(1)
        assert b.frame_counters()["bvh_refits"] == 0
        hb, ob, lb_ = b.trace_closest(o, d), b.trace_any(o, d, tmax), b.read_lights()
    assert np.array_equal(ha, hb) and np.array_equal(oa, ob) and (ha["t"] > 0).mean() > 0.3
    assert np.array_equal(la[0], lb_[0]) and np.array_equal(la[1], lb_[1])
