"""Generates the NanoVDB golden data with the REFERENCE's own vendored NanoVDB (compiled in place into oracle/_ref/ref_nanovdb by
oracle/Makefile; needs /root/reference, so it runs in the build container only):
  tests/golden/nanovdb/{fog5_raw,fog12_zip,ls10_zip}.vndb   small files written by nanovdb::createFogVolumeSphere / createLevelSetSphere
                                                            + io::writeGrid (codec NONE / ZIP)
  tests/golden/nanovdb_reference.npz                        per file (and for the reference's Sandbox/assets/volume/Sphere.vndb, which is
                                                            not copied): meta data, 2000 sampled voxels (value + active state) read through
                                                            nanovdb::ReadAccessor, and sum / bit-fold of every voxel of the index bbox
Usage: python tests/golden/make_golden_nanovdb.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nanovdb_tools as nt   # noqa: E402

TOOL = os.path.join(ROOT, "oracle", "_ref", "ref_nanovdb")
SPHERE = "/root/reference/Lumen_Engine/Sandbox/assets/volume/Sphere.vndb"
FIXTURES = {   # name: kind radius voxel halfwidth cx cy cz [codec]
    "fog5_raw": ["fog", "5", "1", "3", "-40", "-40", "-40"],                 # one upper node at a negative origin, uncompressed
    "fog12_zip": ["fog", "12", "1", "3", "30", "-20", "-4100", "zip"],       # straddles the z = -4096 root-tile boundary
    "ls10_zip": ["ls", "10", "0.5", "3", "21", "12", "20", "zip"],           # narrow-band level set, voxel size 0.5, interior tiles
}


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    out_dir = os.path.join(HERE, "nanovdb")
    os.makedirs(out_dir, exist_ok=True)
    gold = {}
    files = {}
    for name, args in FIXTURES.items():
        path = os.path.join(out_dir, name + ".vndb")
        codec = args[7:]
        subprocess.check_call([TOOL, "make"] + args[:7] + [path] + codec)
        files[name] = path
    files["sphere_asset"] = SPHERE
    with tempfile.TemporaryDirectory() as tmp:
        for name, path in files.items():
            dump = os.path.join(tmp, name + ".bin")
            subprocess.check_call([TOOL, "dump", path, "2000", "7", dump])
            for k, v in nt.read_reference_dump(dump).items():
                gold[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "nanovdb_reference.npz"), **gold)
    print("wrote", sorted(files), "->", os.path.join(HERE, "nanovdb_reference.npz"))


if __name__ == "__main__":
    main()
