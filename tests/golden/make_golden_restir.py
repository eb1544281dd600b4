"""Known answers of the reference's ReSTIR data structures (Reservoir::Update / UpdateWeight, CDF::Insert / Get / BinarySearch,
LumenPT/src/Shaders/CppCommon/ReSTIRData.h:107-302), produced by the reference's OWN header compiled for the host in place
(oracle/ref_shim/ref_restir.cpp -> oracle/_ref/libref_restir.so; needs /root/reference, so this runs in the build container only).
Writes tests/golden/restir_reference.npz. Usage: python tests/golden/make_golden_restir.py"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def reservoir_cases(rng):
    """Update sequences: plain RIS streams, zero weights (geometrically rejected candidates), a leading zero, denormal and huge weights,
    seeds that draw exactly 0 (xorshift state 0) and near 1."""
    cases = []
    for n in (1, 2, 5, 32, 32, 32, 64, 200):
        w = rng.random(n).astype(np.float32) * rng.choice([1e-3, 1.0, 50.0])
        w[rng.random(n) < 0.4] = 0.0
        cases.append((w, rng.integers(1, 2 ** 32, n, dtype=np.uint64).astype(np.uint32), (rng.random(n).astype(np.float32) + 1e-3)))
    cases.append((np.zeros(8, np.float32), rng.integers(1, 2 ** 32, 8, dtype=np.uint64).astype(np.uint32), np.ones(8, np.float32)))          # nothing contributes
    cases.append((np.array([0, 0, 3, 0, 1e-30, 7], np.float32), np.array([5, 0, 0, 9, 11, 0xFFFFFFFF], np.uint32), np.array([1, 1, 0, 1, 1, 1e-12], np.float32)))
    cases.append((np.array([1e30, 1e30, 1, 1e-38], np.float32), np.array([1, 2, 3, 4], np.uint32), np.array([1e-9, 0.5, 2, 3], np.float32)))
    return cases


def cdf_cases(rng):
    cases = []
    for n in (1, 2, 3, 17, 256, 1152, 5000):
        w = (rng.random(n).astype(np.float32) ** 3 * 100 + 1e-3).astype(np.float32)
        v = rng.random(4000).astype(np.float32)
        v[:6] = [0.0, 1.0, 0.5, 1e-8, 0.99999994, 0.25]
        cases.append((w, v))
    # look-ups that land exactly on accumulated sums (interval boundaries): small integer weights keep every partial sum exact
    w = rng.integers(1, 9, 64).astype(np.float32)
    acc = np.cumsum(w.astype(np.float64)); total = acc[-1]
    v = np.concatenate([(acc / total).astype(np.float32), rng.random(500).astype(np.float32)])
    cases.append((w, v))
    return cases


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_restir.so"))
    rng = np.random.default_rng(20261017)
    gold = {}
    for k, (w, seeds, pdfs) in enumerate(reservoir_cases(rng)):
        out, sel = np.zeros(5, np.float32), np.zeros(len(w), np.uint8)
        lib.ref_kat_reservoir(w.ctypes.data_as(C.c_void_p), seeds.ctypes.data_as(C.c_void_p), pdfs.ctypes.data_as(C.c_void_p), C.c_uint(len(w)), out.ctypes.data_as(C.c_void_p), sel.ctypes.data_as(C.c_void_p))
        gold.update({f"res{k}/weights": w, f"res{k}/seeds": seeds, f"res{k}/pdfs": pdfs, f"res{k}/out": out, f"res{k}/selected": sel})
    gold["res_count"] = np.int32(k + 1)
    for k, (w, v) in enumerate(cdf_cases(rng)):
        cdf, idx, pdf = np.zeros(len(w), np.float32), np.zeros(len(v), np.uint32), np.zeros(len(v), np.float32)
        lib.ref_kat_cdf(w.ctypes.data_as(C.c_void_p), C.c_uint(len(w)), v.ctypes.data_as(C.c_void_p), C.c_uint(len(v)), cdf.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), pdf.ctypes.data_as(C.c_void_p))
        gold.update({f"cdf{k}/weights": w, f"cdf{k}/values": v, f"cdf{k}/cdf": cdf, f"cdf{k}/index": idx, f"cdf{k}/pdf": pdf})
    gold["cdf_count"] = np.int32(k + 1)
    np.savez_compressed(os.path.join(HERE, "restir_reference.npz"), **gold)
    print("wrote", os.path.join(HERE, "restir_reference.npz"))


if __name__ == "__main__":
    main()
