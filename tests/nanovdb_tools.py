"""TEST INFRASTRUCTURE — numpy restatement of the NanoVDB (ABI 29) file and grid format, the checker of csrc/lb_nanovdb.cpp.

Follows the NanoVDB the reference vendors (paths under /root/reference/Lumen_Engine/LumenPT/vendor/openvdb/nanovdb/nanovdb/):
  util/IO.h:107-160   segment header + 160-byte grid meta data + name;  :301-352 codecs (ZIP = u64 size + zlib stream)
  NanoVDB.h:1890-1905 GridData, :1775-1784 Map, :2184-2190 TreeData, :2394-2456 RootData + tiles (key = 3 x 21 bits of origin >> 12),
  :2733-2766 InternalData (child = (this + mOffset) as ChildT* + childID), :3022-3040 LeafData<float>.
Pinned against the reference's own library by tests/golden/nanovdb_reference.npz (tests/golden/make_golden_nanovdb.py)."""
import struct
import zlib

import numpy as np

MAGIC = 0x304244566F6E614E
GRID, TREE, ROOT, TILE = 672, 64, 64, 32
UPPER, LOWER, LEAF = 139328, 17472, 2144
UPPER_TABLE, LOWER_TABLE, LEAF_VALUES = 8256, 1088, 96


class Grid:
    def __init__(self, buf: bytes, codec: int, grid_count: int):
        self.buf, self.codec, self.grid_count = buf, codec, grid_count
        g = buf
        assert struct.unpack_from("<Q", g, 0)[0] == MAGIC
        ver, = struct.unpack_from("<I", g, 16)
        self.version = (ver >> 21, (ver >> 10) & 0x7FF, ver & 0x3FF)
        self.name = g[32:288].split(b"\0")[0].decode()
        self.map_matrix = np.array(struct.unpack_from("<9d", g, 288 + 88)).reshape(3, 3)
        self.map_translation = np.array(struct.unpack_from("<3d", g, 288 + 88 + 144))
        self.world_min, self.world_max = np.array(struct.unpack_from("<3d", g, 552)), np.array(struct.unpack_from("<3d", g, 576))
        self.voxel_size = np.array(struct.unpack_from("<3d", g, 600))
        self.grid_class, self.grid_type = struct.unpack_from("<II", g, 624)
        offs = struct.unpack_from("<4Q", g, GRID)
        self.node_count = struct.unpack_from("<4I", g, GRID + 32)
        self.leaves, self.lowers, self.uppers, self.root = (GRID + o for o in offs)
        r = self.root
        bb = struct.unpack_from("<6i", g, r)
        self.index_min, self.index_max = np.array(bb[:3], np.int32), np.array(bb[3:], np.int32)
        self.active_voxels, ntiles = struct.unpack_from("<QI", g, r + 24)
        self.background, self.value_min, self.value_max = struct.unpack_from("<3f", g, r + 36)
        self.tiles = []
        for k in range(ntiles):
            key, child, state, value = struct.unpack_from("<QiIf", g, r + ROOT + TILE * k)
            org = [np.int32(np.uint32((((key >> s) & 0x1FFFFF) << 12) & 0xFFFFFFFF).astype(np.int32)) for s in (42, 21, 0)]
            self.tiles.append((tuple(int(o) for o in org), child, state, value))

    def _child(self, node, node_bytes, table, n, child_bytes):
        off, = struct.unpack_from("<i", self.buf, node + 24)
        cid, = struct.unpack_from("<I", self.buf, node + table + 4 * n)
        return node + off * node_bytes + cid * child_bytes

    def _bit(self, base, n):
        return (self.buf[base + (n >> 3)] >> (n & 7)) & 1

    def value(self, i, j, k):
        """(value, active) of one voxel: ReadAccessor::getValue / isActive."""
        f = lambda off: struct.unpack_from("<f", self.buf, off)[0]
        key = (i & ~4095, j & ~4095, k & ~4095)
        for org, child, state, value in self.tiles:
            if org != key:
                continue
            if child < 0:
                return np.float32(value), bool(state)
            up = self.uppers + child * UPPER
            n = (((i & 4095) >> 7) << 10) + (((j & 4095) >> 7) << 5) + ((k & 4095) >> 7)
            if not self._bit(up + 32 + 4096, n):
                return np.float32(f(up + UPPER_TABLE + 4 * n)), bool(self._bit(up + 32, n))
            lo = self._child(up, UPPER, UPPER_TABLE, n, LOWER)
            n = (((i & 127) >> 3) << 8) + (((j & 127) >> 3) << 4) + ((k & 127) >> 3)
            if not self._bit(lo + 32 + 512, n):
                return np.float32(f(lo + LOWER_TABLE + 4 * n)), bool(self._bit(lo + 32, n))
            lf = self._child(lo, LOWER, LOWER_TABLE, n, LEAF)
            n = ((i & 7) << 6) + ((j & 7) << 3) + (k & 7)
            return np.float32(f(lf + LEAF_VALUES + 4 * n)), bool(self._bit(lf + 16, n))
        return np.float32(self.background), False

    def dense(self):
        """[nz, ny, nx] float32 box index_min..index_max of stored values (tiles and background included)."""
        lo, hi = self.index_min.astype(np.int64), self.index_max.astype(np.int64)
        if np.any(hi < lo):
            return np.zeros((0, 0, 0), np.float32)
        out = np.full(tuple((hi - lo + 1)[::-1]), self.background, np.float32)

        def put(org, block):                      # block indexed [x, y, z]
            a = np.maximum(org, lo); b = np.minimum(org + np.array(block.shape) - 1, hi)
            if np.any(a > b):
                return
            s = block[a[0] - org[0]:b[0] - org[0] + 1, a[1] - org[1]:b[1] - org[1] + 1, a[2] - org[2]:b[2] - org[2] + 1]
            out[a[2] - lo[2]:b[2] - lo[2] + 1, a[1] - lo[1]:b[1] - lo[1] + 1, a[0] - lo[0]:b[0] - lo[0] + 1] = s.transpose(2, 1, 0)

        def mask(base, nbits):
            return np.unpackbits(np.frombuffer(self.buf, np.uint8, nbits // 8, base), bitorder="little").astype(bool)

        for org, child, state, value in self.tiles:
            org = np.array(org, np.int64)
            if child < 0:
                a = np.maximum(org, lo); b = np.minimum(org + 4095, hi)
                if np.all(a <= b):
                    out[a[2] - lo[2]:b[2] - lo[2] + 1, a[1] - lo[1]:b[1] - lo[1] + 1, a[0] - lo[0]:b[0] - lo[0] + 1] = value
                continue
            up = self.uppers + child * UPPER
            ucm = mask(up + 32 + 4096, 32768); utab = np.frombuffer(self.buf, np.float32, 32768, up + UPPER_TABLE)
            for n in range(32768):
                o1 = org + np.array([n >> 10, (n >> 5) & 31, n & 31]) * 128
                if np.any(o1 > hi) or np.any(o1 + 127 < lo):
                    continue
                if not ucm[n]:
                    put(o1, np.full((128, 128, 128), utab[n], np.float32)); continue
                lw = self._child(up, UPPER, UPPER_TABLE, n, LOWER)
                lcm = mask(lw + 32 + 512, 4096); ltab = np.frombuffer(self.buf, np.float32, 4096, lw + LOWER_TABLE)
                for m in range(4096):
                    o0 = o1 + np.array([m >> 8, (m >> 4) & 15, m & 15]) * 8
                    if np.any(o0 > hi) or np.any(o0 + 7 < lo):
                        continue
                    if not lcm[m]:
                        put(o0, np.full((8, 8, 8), ltab[m], np.float32)); continue
                    lf = self._child(lw, LOWER, LOWER_TABLE, m, LEAF)
                    put(o0, np.frombuffer(self.buf, np.float32, 512, lf + LEAF_VALUES).reshape(8, 8, 8))
        return out

    def density(self):
        d = self.dense()
        if self.grid_class == 1:                                  # level set -> fog (sdfToFogVolume ramp)
            with np.errstate(divide="ignore", invalid="ignore"):
                ramp = np.minimum(np.float32(1.0), (-d) / np.float32(self.background)).astype(np.float32)
            return np.where((d < 0) & (self.background > 0), ramp, np.float32(0.0)).astype(np.float32)
        return np.where(d > 0, d, np.float32(0.0)).astype(np.float32)

    def volume_box(self):
        """Object-space box of the stored voxels (voxel i covers [i, i + 1) in index space), as float32."""
        s = np.diag(self.map_matrix)
        w0 = s * self.index_min.astype(np.float64) + self.map_translation
        w1 = s * (self.index_max.astype(np.float64) + 1.0) + self.map_translation
        return np.minimum(w0, w1).astype(np.float32), np.maximum(w0, w1).astype(np.float32)


def read_grid(data: bytes, index: int = 0) -> Grid:
    """nanovdb::io::readGrid(stream, n) (util/IO.h:573-592)."""
    pos, counter, found = 0, 0, None
    while pos + 16 <= len(data):
        magic, ver, count, codec = struct.unpack_from("<QIHH", data, pos)
        if magic != MAGIC:
            raise ValueError("not a NanoVDB file")
        if ver >> 21 != 29:
            raise ValueError("NanoVDB ABI %d" % (ver >> 21))
        pos += 16
        metas = []
        for _ in range(count):
            grid_size, file_size = struct.unpack_from("<QQ", data, pos)
            name_size, = struct.unpack_from("<I", data, pos + 136)
            metas.append((grid_size, file_size)); pos += 160 + name_size
        for grid_size, file_size in metas:
            if counter == index and found is None:
                blob = data[pos:pos + file_size]
                if codec == 1:
                    zsize, = struct.unpack_from("<Q", blob, 0)
                    blob = zlib.decompress(blob[8:8 + zsize])
                elif codec != 0:
                    raise ValueError("codec %d" % codec)
                assert len(blob) == grid_size
                found = (blob, codec)
            pos += file_size; counter += 1
    if found is None:
        raise ValueError("grid index exceeds grid count")
    return Grid(found[0], found[1], counter)


def read_reference_dump(path: str) -> dict:
    """Parses the record oracle/_ref/ref_nanovdb `dump` writes (layout in oracle/ref_shim/ref_nanovdb.cpp)."""
    b = open(path, "rb").read(); o = 0

    def take(fmt):
        nonlocal o
        v = struct.unpack_from("<" + fmt, b, o); o += struct.calcsize("<" + fmt)
        return v
    d = {}
    d["grid_type"], d["grid_class"] = take("II")
    d["index_bbox"] = np.array(take("6i"), np.int32); d["world_bbox"] = np.array(take("6d")); d["voxel_size"] = np.array(take("3d"))
    d["map_matrix"] = np.array(take("9d")); d["map_translation"] = np.array(take("3d"))
    d["active_voxels"] = np.uint64(take("Q")[0]); d["background_min_max"] = np.array(take("3f"), np.float32); d["node_count"] = np.array(take("4I"), np.uint32)
    ns, = take("I")
    rec = np.frombuffer(b, np.dtype([("ijk", "<i4", 3), ("v", "<f4"), ("on", "<u4")]), ns, o); o += 20 * ns
    d["sample_ijk"], d["sample_value"], d["sample_active"] = rec["ijk"].copy(), rec["v"].copy(), (rec["on"] != 0)
    n, = take("Q"); s, = take("d"); f, = take("Q")
    d["dense_count"], d["dense_sum"], d["dense_fold"] = np.uint64(n), np.float64(s), np.uint64(f)
    return d


def fold_bits(dense: np.ndarray) -> int:
    """The rotate-xor fold of all float bit patterns ref_nanovdb computes (x fastest)."""
    f = 0
    for w in dense.reshape(-1).view(np.uint32).tolist():
        f = (((f << 7) | (f >> 57)) & 0xFFFFFFFFFFFFFFFF) ^ w
    return f
