"""Golden data from the stb_image the reference vendors and decodes every texture with (compiled in place by oracle/Makefile ->
oracle/_ref/ref_stb; needs /root/reference, so this runs in the build container only):
  jpeg_reference.npz  every JPEG file under the reference's Sandbox/assets/models (81 files; 65 of them Sponza's textures) -> width, height,
                      SOF type, sampling, and the SHA-256 of the RGBA8 pixels stbi_load(..., 4) returns;
  jpeg_cases.npz      small synthetic files written with Pillow that cover what the shipped ones do not — progressive scans, 4:4:4 / 4:2:2 /
                      4:2:0 / 4:4:0 / 4:1:1 sampling, grey, odd sizes (partial MCUs), restart intervals, Adobe CMYK, 16-bit quantisation
                      tables (quality 1), one-pixel images — each with the pixels stb_image decodes it to (the file rides along, so the
                      test needs no reference tree).
Usage: python tests/golden/make_golden_jpeg.py"""
import hashlib
import io
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ASSETS = "/root/reference/Lumen_Engine/Sandbox/assets/models"


def sof_info(data):
    i = 2
    while i + 9 < len(data):
        if data[i] != 0xFF:
            i += 1; continue
        m = data[i + 1]
        if m in (0xC0, 0xC1, 0xC2):
            n = data[i + 9]
            return m, "+".join(f"{data[i + 11 + 3 * k] >> 4}x{data[i + 11 + 3 * k] & 15}" for k in range(n))
        if m in (0xD8, 0x01) or 0xD0 <= m <= 0xD7:
            i += 2; continue
        i += 2 + (data[i + 2] << 8 | data[i + 3])
    return 0, ""


def stb(tool, path, out):
    res = subprocess.run([tool, path, out], capture_output=True, text=True)
    if res.returncode != 0:
        return None
    w, h = (int(x) for x in res.stdout.split())
    return np.fromfile(out, np.uint8).reshape(h, w, 4)


def synthetic_cases():
    from PIL import Image
    rng = np.random.default_rng(5)

    def picture(w, h):
        y, x = np.mgrid[0:h, 0:w]
        base = np.stack([128 + 100 * np.sin(x / 3.1) * np.cos(y / 4.3), 128 + 90 * np.cos((x + y) / 5.7), 255 * ((x // 4 + y // 4) % 2)], -1)
        return np.clip(base + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)
    cases = {}

    def add(name, img, **kw):
        buf = io.BytesIO(); img.save(buf, "JPEG", **kw); cases[name] = buf.getvalue()
    for sub, tag in ((0, "444"), (1, "422"), (2, "420")):
        add(f"base_{tag}", Image.fromarray(picture(67, 45)), quality=85, subsampling=sub)
        add(f"prog_{tag}", Image.fromarray(picture(70, 51)), quality=75, subsampling=sub, progressive=True)
    add("grey_base", Image.fromarray(picture(33, 17)[..., 0], "L"), quality=90)
    add("grey_prog", Image.fromarray(picture(40, 40)[..., 1], "L"), quality=60, progressive=True)
    add("q1_16bit_tables", Image.fromarray(picture(48, 32)), quality=1)
    add("q100", Image.fromarray(picture(31, 29)), quality=100, subsampling=2)
    add("one_pixel", Image.fromarray(picture(1, 1)), quality=90)
    add("one_row", Image.fromarray(picture(19, 1)), quality=90, subsampling=2)
    add("one_column_prog", Image.fromarray(picture(1, 23)), quality=90, subsampling=2, progressive=True)
    add("restart_base", Image.fromarray(picture(96, 64)), quality=80, subsampling=2, restart_marker_blocks=3)
    add("restart_prog", Image.fromarray(picture(64, 48)), quality=80, subsampling=1, progressive=True, restart_marker_rows=1)
    add("optimised_huffman", Image.fromarray(picture(80, 60)), quality=70, optimize=True)
    add("cmyk_adobe", Image.fromarray(picture(40, 24)).convert("CMYK"), quality=85)
    add("large_420", Image.fromarray(picture(257, 131)), quality=50, subsampling=2)
    # sampling ratios Pillow does not write (4:1:1 = 4x1, 4:4:0 = 1x2: the pixel-replication and the vertical-only up-sampling paths) via OpenCV
    import cv2
    for flag, tag in ((cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411, "411"), (cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, "440")):
        for prog in (0, 1):
            ok, buf = cv2.imencode(".jpg", picture(53 + 11 * prog, 37 + 5 * prog)[..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, flag, cv2.IMWRITE_JPEG_PROGRESSIVE, prog])
            assert ok
            cases[("prog_" if prog else "base_") + tag] = buf.tobytes()
    return cases


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_stb")
    table = []
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "px.rgba")
        for dirpath, _, files in sorted(os.walk(ASSETS)):
            for f in sorted(files):
                if not f.lower().endswith((".jpg", ".jpeg")):
                    continue
                path = os.path.join(dirpath, f)
                px = stb(tool, path, out)
                assert px is not None, path
                m, samp = sof_info(open(path, "rb").read())
                table.append([os.path.relpath(path, ASSETS), str(px.shape[1]), str(px.shape[0]), hex(m), samp, hashlib.sha256(px.tobytes()).hexdigest()])
        np.savez_compressed(os.path.join(HERE, "jpeg_reference.npz"), table=np.array(table))
        print(len(table), "shipped files;", sorted({(r[3], r[4]) for r in table}))
        cases, blob = synthetic_cases(), {}
        for name, data in cases.items():
            path = os.path.join(tmp, name + ".jpg"); open(path, "wb").write(data)
            px = stb(tool, path, out)
            assert px is not None, name
            blob[name + "/file"] = np.frombuffer(data, np.uint8); blob[name + "/rgba"] = px
            print(name, px.shape, sof_info(data))
        np.savez_compressed(os.path.join(HERE, "jpeg_cases.npz"), **blob)


if __name__ == "__main__":
    main()
