#!/usr/bin/env python
"""Generates tests/golden/cornell_gltf.npz from the reference's own CornellBox asset (BASELINE config C1), read where it lies:

    python tests/golden/make_golden_gltf.py

The file /root/reference/Lumen_Engine/Sandbox/assets/models/CornellBox/scene.gltf is interpreted with the numpy/json restatement of the
reference's converter (tests/gltf_tools.py: accessor extraction, default-UV tangent generation, node hierarchy, material mapping) and the
derived arrays are stored: per mesh positions / normals / uvs / tangents / indices / material, per instance mesh + row-major world matrix,
per material base colour / emission / metallic / roughness. The fixture (a) pins the product's C++ glTF loader on machines that do not have
the asset, (b) lets `scenes.cornell_box_reference()` rebuild the reference's exact C1 geometry on the GPU box."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import gltf_tools as gt

SRC = "/root/reference/Lumen_Engine/Sandbox/assets/models/CornellBox/scene.gltf"


def main():
    ref = gt.load_reference_semantics(SRC)
    out = {"num_meshes": np.int32(len(ref["meshes"])), "num_instances": np.int32(len(ref["instances"])), "num_materials": np.int32(len(ref["materials"]))}
    for mi, prims in enumerate(ref["meshes"]):
        assert len(prims) == 1
        for k in ("positions", "normals", "uvs", "tangents", "indices"):
            out[f"mesh{mi}_{k}"] = prims[0][k]
        out[f"mesh{mi}_material"] = np.int32(prims[0]["material"])
    for ii, inst in enumerate(ref["instances"]):
        out[f"inst{ii}_mesh"] = np.int32(inst["mesh"]); out[f"inst{ii}_transform"] = inst["transform"].astype(np.float32)
    for i, m in enumerate(ref["materials"]):
        out[f"mat{i}_color"] = np.asarray(m["diffuse_color"], np.float32); out[f"mat{i}_emission"] = np.asarray(m["emission"], np.float32)
        out[f"mat{i}_metallic_roughness"] = np.asarray([m["metallic_factor"], m["roughness_factor"]], np.float32)
    np.savez_compressed(os.path.join(HERE, "cornell_gltf.npz"), **out)
    print("wrote cornell_gltf.npz:", sum(len(p[0]["indices"]) // 3 for p in ref["meshes"]), "triangles,", len(ref["instances"]), "instances")


if __name__ == "__main__":
    main()
