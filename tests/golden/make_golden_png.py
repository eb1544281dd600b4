"""Golden data from the stb_image the reference vendors and decodes every texture with (compiled in place by oracle/Makefile ->
oracle/_ref/ref_stb; needs /root/reference, so this runs in the build container only): every PNG file under the reference's
Sandbox/assets/models (67 files: RGB, RGBA, palette images of depth 1, 2, 4 and 8) -> width, height and the SHA-256 of the RGBA8 pixels
stbi_load(..., 4) returns. Writes tests/golden/png_reference.npz."""
import hashlib
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ASSETS = "/root/reference/Lumen_Engine/Sandbox/assets/models"


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_stb")
    table = []
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "px.rgba")
        for dirpath, _, files in sorted(os.walk(ASSETS)):
            for f in sorted(files):
                if not f.lower().endswith(".png"):
                    continue
                path = os.path.join(dirpath, f)
                head = open(path, "rb").read(29)
                res = subprocess.run([tool, path, out], capture_output=True, text=True)
                assert res.returncode == 0, path
                w, h = res.stdout.split()
                table.append([os.path.relpath(path, ASSETS), w, h, str(head[24]), str(head[25]), hashlib.sha256(open(out, "rb").read()).hexdigest()])
                print(table[-1])
    np.savez_compressed(os.path.join(HERE, "png_reference.npz"), table=np.array(table))


if __name__ == "__main__":
    main()
