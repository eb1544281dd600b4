"""Known answers of the reference's radiance-deciding device functions, produced by the reference's OWN code compiled for the host in
place (oracle/ref_shim/ref_kernels.cpp -> oracle/_ref/libref_kernels.so; needs /root/reference, so this runs in the build container only):
  Resample, CombineBiased, CombineUnbiased   LumenPT/src/CUDAKernels/ReSTIRKernels.cu:1123-1325
  ShadeIndirect                              LumenPT/src/CUDAKernels/WaveFrontKernels/GPUShadeIndirect.cu:7-146
  ShadeDirect                                LumenPT/src/CUDAKernels/WaveFrontKernels/GPUShadeDirect.cu:42-153
Writes tests/golden/kernels_reference.npz (inputs + the reference's outputs). Usage: python tests/golden/make_golden_kernels.py

Flat layouts (shared with the oracle's lo_kat_* taps): surf44 = position3, normal3, tangent3, incoming3, transport3, t, flags, pad3, mat24;
sample14 = radiance3, normal3, position3, area, contribution3, pdf; reservoir17 = weightSum, sampleCount, weight, sample14;
ray11 = px, py, origin3, direction3, contribution3; shadow12 = px, py, origin3, direction3, tmax, radiance3."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_golden import materials, unit      # the 24 golden materials of the BSDF pin


def surfaces(rng, mats, n, flags_mix=True):
    """n surfaces around the origin: random frames, incoming directions on the upper and (a few) the lower side of the normal, grazing ones,
    every material in turn; some flagged emissive / alpha / miss."""
    s = np.zeros((n, 44), np.float32)
    nrm = unit(rng, n); tan = np.cross(nrm, unit(rng, n)); tan /= np.linalg.norm(tan, axis=1, keepdims=True)
    inc = unit(rng, n)
    flip = (np.sum(inc * nrm, axis=1) > 0) & (rng.random(n) < 0.9)          # 90 %: arriving from above the surface
    inc[flip] *= -1
    graz = rng.random(n) < 0.03                                              # near-perpendicular to the normal: the 3 * EPSILON rejection
    inc[graz] = (tan[graz] + nrm[graz] * (rng.random((int(graz.sum()), 1)) - 0.5) * 4e-4).astype(np.float32)
    inc /= np.linalg.norm(inc, axis=1, keepdims=True)
    s[:, 0:3] = rng.uniform(-2, 2, (n, 3)); s[:, 3:6] = nrm; s[:, 6:9] = tan; s[:, 9:12] = inc
    s[:, 12:15] = rng.random((n, 3)) * rng.choice([1.0, 0.2, 3.0], (n, 1)); s[:, 15] = rng.uniform(0.1, 30, n)
    if flags_mix:
        f = rng.random(n); s[:, 16] = np.where(f < 0.04, 1, np.where(f < 0.09, 2, np.where(f < 0.12, 4, 0)))
    s[:, 20:44] = mats[np.arange(n) % mats.shape[0]]
    return s


def light_samples(rng, surf, n):
    """n light samples around a surface: most above its horizon and facing it, some behind, some facing away, one closer than 1 cm."""
    p = surf[0:3]; nrm = surf[3:6]
    out = np.zeros((n, 14), np.float32)
    d = unit(rng, n); d[np.sum(d * nrm, axis=1) < 0] *= -1
    d[rng.random(n) < 0.1] *= -1                                             # below the horizon
    dist = rng.uniform(0.3, 12, n).astype(np.float32); dist[0] = 0.005
    out[:, 6:9] = p + d * dist[:, None]
    ln = -d + unit(rng, n) * 0.7; ln /= np.linalg.norm(ln, axis=1, keepdims=True)
    ln[rng.random(n) < 0.1] *= -1                                            # facing away
    out[:, 3:6] = ln; out[:, 0:3] = rng.random((n, 3)) * 200; out[:, 9] = rng.uniform(0.001, 0.5, n)
    out[:, 10:13] = rng.random((n, 3)); out[:, 13] = rng.random(n)
    return out


def frame_surfaces(rng, mats, W, H, eye):
    """Primary-hit records of a camera looking at a floor (y = 0) and a back wall (z = -6): smooth depth, two orientations with a perturbed
    shading normal (a few beyond the 25-degree similarity bound), materials in 8x8-pixel blocks, 3 % flagged pixels."""
    ys, xs = np.mgrid[0:H, 0:W]
    d = np.stack([(xs / W - 0.5) * 2.6, -(ys / H - 0.5) * 2.0 - 0.35, -np.ones_like(xs, float)], -1)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    t_floor = np.where(d[..., 1] < -1e-6, -eye[1] / np.minimum(d[..., 1], -1e-6), 1e30)
    t_wall = (-6.0 - eye[2]) / d[..., 2]
    t = np.minimum(t_floor, t_wall); on_floor = t_floor < t_wall
    pos = eye + d * t[..., None]
    nrm = np.where(on_floor[..., None], np.array([0.0, 1.0, 0.0]), np.array([0.0, 0.0, 1.0]))
    bend = rng.normal(size=(H, W, 3)) * np.where(rng.random((H, W, 1)) < 0.06, 0.9, 0.05)
    nrm = nrm + bend; nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    tan = np.cross(nrm, np.array([0.3, 0.5, 0.8])); tan /= np.linalg.norm(tan, axis=-1, keepdims=True)
    s = np.zeros((H, W, 44), np.float32)
    s[..., 0:3] = pos; s[..., 3:6] = nrm; s[..., 6:9] = tan; s[..., 9:12] = d; s[..., 12:15] = 1.0; s[..., 15] = t
    f = rng.random((H, W)); s[..., 16] = np.where(f < 0.01, 1, np.where(f < 0.02, 2, np.where(f < 0.03, 4, 0)))
    s[..., 20:44] = mats[((ys // 8) * 7 + (xs // 8) * 3) % mats.shape[0]]
    return s.reshape(H * W, 44)


def main():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_kernels.so"))
    P = C.c_void_p
    ref.ref_kat_resample.argtypes = [P, C.c_uint, P, P]
    ref.ref_kat_combine.argtypes = [P, C.c_uint, P, P, C.c_uint, C.c_int, P]
    ref.ref_kat_shade_indirect.argtypes = [P, C.c_uint, C.c_uint, C.c_uint, P, C.c_uint]; ref.ref_kat_shade_indirect.restype = C.c_uint
    ref.ref_kat_shade_direct.argtypes = [P, C.c_uint, C.c_uint, C.c_uint, P, P, C.c_uint, P, C.c_uint]; ref.ref_kat_shade_direct.restype = C.c_uint
    ref.ref_kat_ris.argtypes = [P, C.c_uint, C.c_uint, C.c_uint, C.c_uint, P, P, C.c_uint, P, P, P]
    ref.ref_kat_visibility_rays.argtypes = [P, P, C.c_uint, C.c_uint, P]; ref.ref_kat_visibility_rays.restype = C.c_uint
    ref.ref_kat_temporal.argtypes = [P, P, P, P, P, C.c_uint, C.c_uint, C.c_uint, P, P]
    ref.ref_kat_spatial.argtypes = [P, P, C.c_uint, C.c_uint, C.c_uint, P]
    ref.ref_kat_combine_buffers.argtypes = [P, P, P, C.c_uint, C.c_uint, C.c_uint]
    rng = np.random.default_rng(20261018)
    mats = materials(np.random.default_rng(7))
    out = {"mats": mats}

    # ---- Resample: every material x 48 light samples
    rs_surf = surfaces(rng, mats, mats.shape[0], flags_mix=False)
    rs_in = np.stack([light_samples(rng, rs_surf[i], 48) for i in range(rs_surf.shape[0])])
    rs_out = np.zeros_like(rs_in)
    for i in range(rs_surf.shape[0]):
        ref.ref_kat_resample(rs_in[i].ctypes.data, 48, rs_surf[i].ctypes.data, rs_out[i].ctypes.data)
    out.update(resample_surf=rs_surf, resample_in=rs_in, resample_out=rs_out)

    # ---- CombineBiased / CombineUnbiased: 2 .. 6 reservoirs per call (2 = temporal / buffer merge, up to 5 = spatial), zero-weight and
    # zero-count reservoirs, seeds incl. the all-zero xorshift state
    cb_surf, cb_res, cb_from, cb_seed, cb_n, cb_out_b, cb_out_u = [], [], [], [], [], [], []
    for case in range(160):
        n = int(rng.integers(2, 7))
        px = surfaces(rng, mats[case % mats.shape[0]][None], 1, flags_mix=False)[0]
        frm = surfaces(rng, mats, n, flags_mix=False)
        frm[:, 0:3] = px[0:3] + rng.normal(size=(n, 3)).astype(np.float32) * 0.05; frm[:, 3:6] = px[3:6]
        res = np.zeros((6, 17), np.float32)
        smp = light_samples(rng, px, n)
        res[:n, 3:] = smp; res[:n, 0] = rng.random(n) * 10; res[:n, 1] = rng.integers(0, 640, n); res[:n, 2] = rng.random(n) * 3
        res[:n, 2][rng.random(n) < 0.2] = 0.0; res[:n, 1][rng.random(n) < 0.1] = 0.0
        seed = np.uint32(0 if case % 40 == 0 else rng.integers(1, 2 ** 32))
        ob, ou = np.zeros(17, np.float32), np.zeros(17, np.float32)
        frm6 = np.zeros((6, 44), np.float32); frm6[:n] = frm
        ref.ref_kat_combine(res.ctypes.data, n, px.ctypes.data, frm6.ctypes.data, int(seed), 0, ob.ctypes.data)
        ref.ref_kat_combine(res.ctypes.data, n, px.ctypes.data, frm6.ctypes.data, int(seed), 1, ou.ctypes.data)
        # a call in which no Update ever selects a sample returns the default-constructed LightSample, whose unshadowedPathContribution the
        # reference leaves UNINITIALISED (LightSample(), ReSTIRData.h:98 — stack garbage here): recorded as zero, which is what the oracle's
        # and the CUDA library's zero-initialised reservoirs hold
        for o in (ob, ou):
            if not o[3:12].any() and o[16] == 0: o[13:16] = 0
        cb_surf.append(px); cb_res.append(res); cb_from.append(frm6); cb_seed.append(seed); cb_n.append(n); cb_out_b.append(ob); cb_out_u.append(ou)
    out.update(combine_surf=np.stack(cb_surf), combine_res=np.stack(cb_res), combine_from=np.stack(cb_from), combine_seed=np.array(cb_seed, np.uint32),
               combine_n=np.array(cb_n, np.uint32), combine_out_biased=np.stack(cb_out_b), combine_out_unbiased=np.stack(cb_out_u))

    # ---- ShadeIndirect: a 48 x 32 grid of surfaces (every material 64 times; flagged, grazing and back-facing ones among them), 3 frame seeds
    W, H = 48, 32
    si_surf = surfaces(rng, mats, W * H)
    si_seeds = np.array([0x9E3779B9, 12345, 0xDEADBEEF], np.uint32)
    si_rays, si_counts = [], []
    for sd in si_seeds:
        rays = np.zeros((W * H, 11), np.float32)
        si_counts.append(ref.ref_kat_shade_indirect(si_surf.ctypes.data, W, H, int(sd), rays.ctypes.data, W * H)); si_rays.append(rays)
    out.update(indirect_surf=si_surf, indirect_seeds=si_seeds, indirect_wh=np.array([W, H], np.uint32), indirect_rays=np.stack(si_rays), indirect_counts=np.array(si_counts, np.uint32))

    # ---- ShadeDirect: the same kind of grid, 200 lights sorted by mean radiance with the CDF built by CDF::Insert
    sd_surf = surfaces(rng, mats, W * H)
    nl = 200
    lights = np.zeros((nl, 16), np.float32)
    p0 = rng.uniform(-6, 6, (nl, 3)); e1 = rng.normal(size=(nl, 3)) * 0.3; e2 = rng.normal(size=(nl, 3)) * 0.3
    lights[:, 0:3] = p0; lights[:, 3:6] = p0 + e1; lights[:, 6:9] = p0 + e2
    nrm = np.cross(e1, e2); area = np.linalg.norm(nrm, axis=1) / 2; nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    lights[:, 9:12] = nrm; lights[:, 12:15] = rng.random((nl, 3)) * rng.choice([1.0, 50.0, 200.0], (nl, 1)); lights[:, 15] = area
    key = ((lights[:, 12] + lights[:, 13] + lights[:, 14]) / np.float32(3.0)).astype(np.float32)
    lights = lights[np.argsort(key, kind="stable")]; key = np.sort(key, kind="stable")
    sd_rays, sd_counts = [], []
    for sd in si_seeds:
        rays = np.zeros((W * H, 12), np.float32)
        sd_counts.append(ref.ref_kat_shade_direct(sd_surf.ctypes.data, W, H, int(sd), lights.ctypes.data, key.ctypes.data, nl, rays.ctypes.data, W * H)); sd_rays.append(rays)
    out.update(direct_surf=sd_surf, direct_lights=lights, direct_cdf_weights=key, direct_rays=np.stack(sd_rays), direct_counts=np.array(sd_counts, np.uint32))

    # ---- the ReSTIR kernels, whole, chained over a 96 x 64 frame as ReSTIR::Run chains them (ReSTIR.cpp:125-220): light bags + RIS on the current
    # and on the previous frame's surfaces, visibility rays, temporal reuse (motion vectors of a few pixels), two spatial iterations, buffer merge
    FW, FH = 96, 64
    lit = lights.copy(); lit[:, [1, 4, 7]] = np.abs(lit[:, [1, 4, 7]]) * 0.5 + 1.0          # the lights above the floor
    e1, e2 = lit[:, 3:6] - lit[:, 0:3], lit[:, 6:9] - lit[:, 0:3]
    nrm = np.cross(e1, e2); lit[:, 15] = np.linalg.norm(nrm, axis=1) / 2; lit[:, 9:12] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    lit[::2, 9:12] *= -1                                                                        # half of them face down
    cur = frame_surfaces(rng, mats, FW, FH, np.array([0.0, 2.0, 4.0])); prev = frame_surfaces(rng, mats, FW, FH, np.array([0.06, 2.0, 4.02]))
    nb = 50 * 1000
    bag_pdf, bag_p0x = np.zeros(nb, np.float32), np.zeros(nb, np.float32)
    res_cur, res_prev = np.zeros((FW * FH, 17), np.float32), np.zeros((FW * FH, 17), np.float32)
    def clean(r17):
        # reservoirs in which no Update ever selected a sample hold the default-constructed LightSample, whose unshadowedPathContribution the
        # reference leaves uninitialised (ReSTIRData.h:98): recorded as zero (see the combine cases above); done before the array is passed on
        never = ~r17[:, 3:12].any(axis=1) & (r17[:, 16] == 0)
        r17[never, 13:16] = 0
    ref.ref_kat_ris(prev.ctypes.data, FW, FH, 0x1234567, 0x89ABCDE, lit.ctypes.data, key.ctypes.data, nl, bag_pdf.ctypes.data, bag_p0x.ctypes.data, res_prev.ctypes.data)
    ref.ref_kat_ris(cur.ctypes.data, FW, FH, 0xA5A5A5A5, 0x0F1E2D3C, lit.ctypes.data, key.ctypes.data, nl, bag_pdf.ctypes.data, bag_p0x.ctypes.data, res_cur.ctypes.data)
    clean(res_prev); clean(res_cur)
    vis = np.zeros((FW * FH, 8), np.float32)
    nvis = ref.ref_kat_visibility_rays(cur.ctypes.data, res_cur.ctypes.data, FW, FH, vis.ctypes.data)
    motion = (rng.integers(-3, 4, (FW * FH, 2)) / np.array([FW, FH])).astype(np.float16).astype(np.float32)
    motion[rng.random(FW * FH) < 0.5] = 0
    tmp_out = res_cur.copy(); direct = np.zeros((FW * FH, 4), np.float32)
    ref.ref_kat_temporal(cur.ctypes.data, prev.ctypes.data, res_cur.ctypes.data, res_prev.ctypes.data, motion.ctypes.data, FW, FH, 0x5EED0001, tmp_out.ctypes.data, direct.ctypes.data)
    clean(tmp_out)
    sp1 = res_prev.copy()                                                                       # stale content of the output buffer (Reset() keeps the sample)
    ref.ref_kat_spatial(cur.ctypes.data, tmp_out.ctypes.data, FW, FH, 0x5EED0002, sp1.ctypes.data)
    clean(sp1)
    sp2 = np.zeros_like(sp1)
    ref.ref_kat_spatial(cur.ctypes.data, sp1.ctypes.data, FW, FH, 0x5EED0002, sp2.ctypes.data)
    clean(sp2)
    merged = tmp_out.copy()
    ref.ref_kat_combine_buffers(cur.ctypes.data, merged.ctypes.data, sp2.ctypes.data, FW, FH, 0x5EED0003)
    clean(merged)
    # the same temporal and spatial kernels with ReSTIRSettings::enableBiased flipped to false (the one constant changed in a build-time overlay of
    # ReSTIRData.h, oracle/Makefile -> libref_kernels_unbiased.so): the CombineUnbiased branches :905-970 and :1116, dead in the shipped build
    unb = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_kernels_unbiased.so"))
    unb.ref_kat_temporal.argtypes = ref.ref_kat_temporal.argtypes; unb.ref_kat_spatial.argtypes = ref.ref_kat_spatial.argtypes
    tmp_unb = res_cur.copy(); direct_unb = np.zeros((FW * FH, 4), np.float32)
    unb.ref_kat_temporal(cur.ctypes.data, prev.ctypes.data, res_cur.ctypes.data, res_prev.ctypes.data, motion.ctypes.data, FW, FH, 0x5EED0001, tmp_unb.ctypes.data, direct_unb.ctypes.data)
    clean(tmp_unb)
    sp_unb = res_prev.copy()
    unb.ref_kat_spatial(cur.ctypes.data, tmp_out.ctypes.data, FW, FH, 0x5EED0002, sp_unb.ctypes.data)
    clean(sp_unb)
    out.update(frame_temporal_unbiased=tmp_unb, frame_spatial_unbiased=sp_unb)
    print("unbiased: temporal differs from biased in", int((tmp_unb != tmp_out).any(axis=1).sum()), "reservoirs, spatial in", int((sp_unb != sp1).any(axis=1).sum()))
    out.update(frame_wh=np.array([FW, FH], np.uint32), frame_lights=lit, frame_cur=cur, frame_prev=prev, frame_bag_pdf=bag_pdf, frame_bag_p0x=bag_p0x,
               frame_res_cur=res_cur, frame_res_prev=res_prev, frame_vis=vis[:nvis], frame_motion=motion, frame_temporal=tmp_out, frame_direct=direct,
               frame_spatial1=sp1, frame_spatial2=sp2, frame_merged=merged)
    print("frame: RIS selected", int((res_cur[:, 2] > 0).sum()), "of", FW * FH, "| visibility rays", nvis, "| temporal changed", int((tmp_out != res_cur).any(axis=1).sum()),
          "| spatial combined", int((sp1[:, 1] > 0).sum()), int((sp2[:, 1] > 0).sum()))

    path = os.path.join(HERE, "kernels_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, "indirect rays", si_counts, "shadow rays", sd_counts)


if __name__ == "__main__":
    main()
