"""Small synthetic PNG files for the corners the reference's shipped textures do not reach — grey images of depth 1, 2, 4, 8, 16 and RGB
images of depth 8, 16 with a tRNS colour key, grey + alpha, palette with tRNS, Adam7 interlacing (incl. sizes with empty passes), all five row filters — decoded by the stb_image
the reference vendors (oracle/_ref/ref_stb; needs /root/reference). Writes tests/golden/png_cases.npz: the files and stb's RGBA8 pixels."""
import os
import struct
import subprocess
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def chunk(kind, body):
    return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)


def filtered_rows(samples, depth, rng):
    """One (sub-)image: rows packed MSB first, each with a random filter type (encoded properly)."""
    h, w, channels = samples.shape
    bits = channels * depth; stride = (w * bits + 7) // 8; bpp = max(1, bits // 8)
    rows = np.zeros((h, stride), np.uint8)
    for y in range(h):
        acc = 0; nb = 0; out = []
        for v in samples[y].reshape(-1):
            acc = (acc << depth) | int(v); nb += depth
            while nb >= 8:
                out.append((acc >> (nb - 8)) & 255); nb -= 8
        if nb:
            out.append((acc << (8 - nb)) & 255)
        rows[y] = out
    raw = bytearray()
    for y in range(h):
        f = int(rng.integers(0, 5)); raw.append(f)
        for x in range(stride):
            a = int(rows[y, x - bpp]) if x >= bpp else 0; b = int(rows[y - 1, x]) if y else 0; c = int(rows[y - 1, x - bpp]) if (y and x >= bpp) else 0
            if f == 4:
                p = a + b - c; pa, pb, pc = abs(p - a), abs(p - b), abs(p - c); pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            else:
                pred = (0, a, b, (a + b) >> 1)[f]
            raw.append((int(rows[y, x]) - pred) & 255)
    return raw


def png(w, h, depth, colour, samples, rng, plte=None, trns=None, interlace=False):
    """samples: (h, w, channels) integers below 2**depth."""
    if interlace:                                             # Adam7: seven sub-images, empty ones left out
        raw = bytearray()
        for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            sub = samples[y0::dy, x0::dx]
            if sub.shape[0] and sub.shape[1]:
                raw += filtered_rows(sub, depth, rng)
    else:
        raw = filtered_rows(samples, depth, rng)
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, colour, 0, 0, int(interlace)))
    if plte is not None:
        data += chunk(b"PLTE", bytes(plte))
    if trns is not None:
        data += chunk(b"tRNS", bytes(trns))
    z = zlib.compress(bytes(raw), 6)
    return data + chunk(b"IDAT", z[:len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2:]) + chunk(b"IEND", b"")


def main():
    rng = np.random.default_rng(68)
    cases = {}
    for depth in (1, 2, 4, 8, 16):
        w, h = 13, 7
        g = rng.integers(0, 2 ** min(depth, 3), (h, w, 1)) if depth > 2 else rng.integers(0, 2 ** depth, (h, w, 1))
        cases[f"grey{depth}_key"] = png(w, h, depth, 0, g, rng, trns=struct.pack(">H", 1))
        cases[f"grey{depth}"] = png(w, h, depth, 0, rng.integers(0, 2 ** depth, (h, w, 1)), rng)
    for depth in (8, 16):
        rgb = rng.integers(0, 3, (6, 9, 3))
        cases[f"rgb{depth}_key"] = png(9, 6, depth, 2, rgb, rng, trns=struct.pack(">3H", 1, 2, 0))
        cases[f"grey_alpha{depth}"] = png(5, 4, depth, 4, rng.integers(0, 2 ** depth, (4, 5, 2)), rng)
        cases[f"rgba{depth}"] = png(5, 4, depth, 6, rng.integers(0, 2 ** depth, (4, 5, 4)), rng)
    cases["rgb16_wide"] = png(7, 3, 16, 2, rng.integers(0, 65536, (3, 7, 3)), rng)
    for depth in (1, 2, 4, 8):
        n = min(2 ** depth, 11)
        cases[f"palette{depth}_trns"] = png(11, 5, depth, 3, rng.integers(0, n, (5, 11, 1)), rng, plte=rng.integers(0, 256, 3 * n).astype(np.uint8), trns=rng.integers(0, 256, max(1, n - 1)).astype(np.uint8))
    # Adam7-interlaced files, including sizes that leave some of the seven passes empty
    for (w, h) in ((1, 1), (3, 2), (5, 9), (17, 11)):
        cases[f"lace_rgba8_{w}x{h}"] = png(w, h, 8, 6, rng.integers(0, 256, (h, w, 4)), rng, interlace=True)
        cases[f"lace_palette2_{w}x{h}"] = png(w, h, 2, 3, rng.integers(0, 4, (h, w, 1)), rng, plte=rng.integers(0, 256, 12).astype(np.uint8), interlace=True)
    cases["lace_grey1_19x10"] = png(19, 10, 1, 0, rng.integers(0, 2, (10, 19, 1)), rng, interlace=True)
    cases["lace_rgb16_9x9_key"] = png(9, 9, 16, 2, rng.integers(0, 2, (9, 9, 3)), rng, trns=struct.pack(">3H", 1, 0, 1), interlace=True)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_stb")
    gold = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, data in cases.items():
            path = os.path.join(tmp, name + ".png"); open(path, "wb").write(data)
            res = subprocess.run([tool, path, path + ".rgba"], capture_output=True, text=True)
            assert res.returncode == 0, name
            w, h = (int(v) for v in res.stdout.split())
            gold[name + "/file"] = np.frombuffer(data, np.uint8)
            gold[name + "/rgba"] = np.fromfile(path + ".rgba", np.uint8).reshape(h, w, 4)
            print(name, w, h, "alpha values", sorted(set(gold[name + "/rgba"][..., 3].reshape(-1).tolist()))[:4])
    np.savez_compressed(os.path.join(HERE, "png_cases.npz"), **gold)


if __name__ == "__main__":
    main()
