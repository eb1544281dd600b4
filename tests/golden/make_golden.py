#!/usr/bin/env python
"""Generates the committed golden fixtures. Run in the build container (needs /root/reference for oracle/_ref):

    make -C oracle && python tests/golden/make_golden.py

bsdf_reference.npz   outputs of the REFERENCE's own BSDF / RNG / material-packing code (disney.cuh, ggxmdf.cuh, frosted.cuh,
                     bsdf_math.cuh, RandomUtilities.cuh, MaterialStructs.h compiled for the host in place -> oracle/_ref/libref_bsdf.so)
                     on seeded inputs. They pin the oracle's restatement (tests/test_oracle_golden.py) and, on the GPU box where
                     /root/reference does not exist, the CUDA BSDF (tests/test_gpu_bsdf.py).
cornell_oracle.npz   the oracle's own C1 Cornell frame (hit ids, t, HDR) — a regression anchor for the restated wavefront, and the
                     fixture the GPU parity test re-checks without re-running the oracle at full size.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


class RefMaterial(C.Structure):
    _fields_ = [("color", C.c_float * 4), ("emissive", C.c_float * 4), ("transmittance", C.c_float * 4), ("tint", C.c_float * 4), ("params", C.c_uint * 4)]


def load_ref():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_bsdf.so"))
    ref.ref_wang_hash.restype = C.c_uint; ref.ref_wang_hash.argtypes = [C.c_uint]
    ref.ref_random_int.restype = C.c_uint; ref.ref_random_int.argtypes = [C.POINTER(C.c_uint)]
    ref.ref_random_float.restype = C.c_float; ref.ref_random_float.argtypes = [C.POINTER(C.c_uint)]
    ref.ref_pack_material.argtypes = [C.POINTER(RefMaterial), C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_float] * 13
    ref.ref_evaluate_bsdf.argtypes = [C.POINTER(RefMaterial)] + [C.c_void_p] * 6
    ref.ref_sample_bsdf.argtypes = [C.POINTER(RefMaterial)] + [C.c_void_p] * 4 + [C.c_float] * 4 + [C.c_void_p] * 4
    return ref


def ref_material(ref, m24):
    m = RefMaterial()
    c, t, ti = (np.ascontiguousarray(m24[a:b], np.float32) for a, b in ((0, 4), (4, 7), (8, 11)))
    # mat24: color4, transmittance3, ior, tint3, luminance, metallic, subsurface, specular, roughness, spectint, anisotropic, sheen, sheentint, clearcoat, clearcoatgloss, transmission, pad
    ref.ref_pack_material(C.byref(m), c.ctypes.data, t.ctypes.data, ti.ctypes.data, *(float(x) for x in (m24[11], m24[7], *m24[12:23])))
    return m


def unit(rng, n):
    v = rng.normal(size=(n, 3)); return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def materials(rng):
    base = dict(color=(0.8, 0.7, 0.6, 1.0), transmittance=(0.1, 0.2, 0.3), ior=1.5, tint=(0.9, 0.8, 0.7), luminance=1.0, metallic=0.0, subsurface=0.0, specular=0.0, roughness=1.0,
                spectint=0.0, anisotropic=0.0, sheen=0.0, sheentint=0.0, clearcoat=0.0, clearcoatgloss=1.0, transmission=0.0)
    variants = [{}, dict(metallic=1.0, roughness=0.3), dict(specular=0.8, roughness=0.4, spectint=0.5), dict(clearcoat=1.0, clearcoatgloss=0.8, roughness=0.7, specular=0.3),
                dict(sheen=0.8, sheentint=0.4, subsurface=0.5, roughness=0.9), dict(transmission=0.9, roughness=0.15, specular=0.5), dict(transmission=1.0, roughness=0.5),
                dict(anisotropic=0.7, metallic=0.8, roughness=0.35, specular=0.6), dict(roughness=0.02, specular=1.0, metallic=0.5), dict(transmission=0.5, ior=1.0, roughness=0.3, specular=0.4),
                dict(subsurface=1.0, roughness=0.6), dict(luminance=0.3, specular=1.0, clearcoat=0.5, sheen=0.5, roughness=0.5)]
    for _ in range(12):
        variants.append({k: float(rng.random()) for k in ("metallic", "subsurface", "specular", "spectint", "anisotropic", "sheen", "sheentint", "clearcoat", "clearcoatgloss", "transmission")}
                        | {"roughness": float(0.02 + 0.98 * rng.random()), "ior": float(1.05 + rng.random())})
    out = []
    for v in variants:
        d = dict(base); d.update(v)
        out.append(np.array([*d["color"], *d["transmittance"], d["ior"], *d["tint"], d["luminance"], d["metallic"], d["subsurface"], d["specular"], d["roughness"], d["spectint"],
                             d["anisotropic"], d["sheen"], d["sheentint"], d["clearcoat"], d["clearcoatgloss"], d["transmission"], 0.0], np.float32))
    return np.stack(out)


def main():
    ref = load_ref()
    rng = np.random.default_rng(20261017)
    mats = materials(rng)
    M, K = mats.shape[0], 160
    ev_in = np.zeros((M, K, 12), np.float32); ev_out = np.zeros((M, K, 4), np.float32)
    sa_in = np.zeros((M, K, 12), np.float32); sa_out = np.zeros((M, K, 8), np.float32)
    packed = np.zeros((M, 4), np.uint32)
    for i in range(M):
        rm = ref_material(ref, mats[i]); packed[i] = list(rm.params)
        n = unit(rng, K); t = np.cross(n, unit(rng, K)); t = (t / np.linalg.norm(t, axis=1, keepdims=True)).astype(np.float32)
        wo, wi = unit(rng, K), unit(rng, K)
        flip = (np.sum(wo * n, axis=1) < 0) & (rng.random(K) < 0.8); wo[flip] *= -1          # mostly front-facing viewers, some back-facing
        flip = (np.sum(wi * n, axis=1) < 0) & (rng.random(K) < 0.6); wi[flip] *= -1
        r012 = rng.random((K, 3)).astype(np.float32)
        ev_in[i] = np.concatenate([n, t, wo, wi], 1); sa_in[i] = np.concatenate([n, t, wo, r012], 1)
        for k in range(K):
            b = np.zeros(3, np.float32); pdf = C.c_float(0)
            a = [np.ascontiguousarray(x) for x in (n[k], t[k], wo[k], wi[k])]
            ref.ref_evaluate_bsdf(C.byref(rm), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, b.ctypes.data, C.addressof(pdf))
            ev_out[i, k] = [*b, pdf.value]
            b = np.zeros(3, np.float32); w = np.zeros(3, np.float32); pdf = C.c_float(0); spec = C.c_int(0)
            ref.ref_sample_bsdf(C.byref(rm), a[0].ctypes.data, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, 1.0, float(r012[k, 0]), float(r012[k, 1]), float(r012[k, 2]),
                                b.ctypes.data, w.ctypes.data, C.addressof(pdf), C.addressof(spec))
            sa_out[i, k] = [*b, *w, pdf.value, float(spec.value)]
    seeds = rng.integers(0, 2**32, 256, dtype=np.uint64).astype(np.uint32)
    hashes = np.array([ref.ref_wang_hash(int(s)) for s in seeds], np.uint32)
    streams = np.zeros((16, 8), np.uint32); floats = np.zeros((16, 8), np.float32)
    for i in range(16):
        s = C.c_uint(int(hashes[i]) | 1)
        for k in range(8):
            streams[i, k] = ref.ref_random_int(C.byref(s))
        s = C.c_uint(int(hashes[i]) | 1)
        for k in range(8):
            floats[i, k] = ref.ref_random_float(C.byref(s))
    np.savez_compressed(os.path.join(HERE, "bsdf_reference.npz"), mats=mats, packed=packed, eval_in=ev_in, eval_out=ev_out, sample_in=sa_in, sample_out=sa_out,
                        seeds=seeds, hashes=hashes, rng_u32=streams, rng_f32=floats)
    print("bsdf_reference.npz:", M, "materials x", K, "directions; NaN eval", int(np.isnan(ev_out).sum()), "NaN sample", int(np.isnan(sa_out).sum()))

    # the oracle's own Cornell C1 frame
    import __graft_entry__ as ge
    from lumenrenderer_b200 import api, scenes
    r = api.Renderer(ge.oracle_bindings(), api.Settings(width=256, height=256, depth=2, restir=False))
    r.load_scene(scenes.cornell_box()); r.render_frames(1)
    hits = r.read_primary_hits(); hdr = r.read_hdr()
    np.savez_compressed(os.path.join(HERE, "cornell_oracle.npz"), instance=hits["instance"].astype(np.uint8), primitive=hits["primitive"].astype(np.uint8), t=hits["t"],
                        hdr=hdr[..., :3].astype(np.float32), counters=np.array(list(r.frame_counters().values()), np.uint64))
    print("cornell_oracle.npz written; mean radiance", float(hdr[..., :3].mean()))


if __name__ == "__main__":
    main()
