"""Test configuration. `-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI export check, gloo multi-process logic.
`-m gpu`: parity tests proper — the CUDA library through its C ABI against the CPU oracle on identical seeded inputs."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_SO = os.path.join(ROOT, "oracle", "liblumen_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_bsdf.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """ctypes bindings of the CPU oracle (built on demand; test infrastructure only)."""
    from lumenrenderer_b200 import api
    if not os.path.exists(ORACLE_SO):
        import subprocess
        subprocess.run(["make", "liblumen_oracle.so"], cwd=os.path.join(ROOT, "oracle"), check=True)
    return api.Bindings(ctypes.CDLL(ORACLE_SO), "lo_")


@pytest.fixture(scope="session")
def gpu():
    """The product: liblumen_b200.so. Fails loudly when it is missing — there is no fallback to test instead."""
    import lumenrenderer_b200 as lr
    return lr.bindings()


def rel_l1(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-30))


def read_png_rgba8(path):
    """Minimal PNG reader for the output-stage tests (8-bit RGBA, non-interlaced; all five filter types; zlib from the standard library)."""
    import struct
    import zlib
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n", "not a PNG"
    pos, idat, w, h, seen_end = 8, b"", 0, 0, False
    while pos < len(data):
        n, = struct.unpack(">I", data[pos:pos + 4]); kind = data[pos + 4:pos + 8]; body = data[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(kind + body) == crc, f"bad CRC in {kind}"
        if kind == b"IHDR":
            w, h, depth, colour, comp, filt, lace = struct.unpack(">IIBBBBB", body)
            assert (depth, colour, comp, filt, lace) == (8, 6, 0, 0, 0)
        elif kind == b"IDAT":
            idat += body
        elif kind == b"IEND":
            seen_end = True
        pos += 12 + n
    assert seen_end
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, w * 4 + 1)
    out = np.zeros((h, w * 4), np.int64)
    for y in range(h):
        f, r = int(raw[y, 0]), raw[y, 1:].astype(np.int64)
        up = out[y - 1] if y else np.zeros(w * 4, np.int64)
        if f == 0:
            out[y] = r
        elif f == 1:
            out[y] = (r.reshape(w, 4).cumsum(0) % 256).reshape(-1)
        elif f == 2:
            out[y] = (r + up) % 256
        else:                                   # Average / Paeth: sequential
            row = np.zeros(w * 4, np.int64)
            for x in range(w * 4):
                a = row[x - 4] if x >= 4 else 0; b = up[x]; c = up[x - 4] if x >= 4 else 0
                if f == 3:
                    pred = (a + b) // 2
                else:
                    p = a + b - c; pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                row[x] = (r[x] + pred) % 256
            out[y] = row
    return out.astype(np.uint8).reshape(h, w, 4)
