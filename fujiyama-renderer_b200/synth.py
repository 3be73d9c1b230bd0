"""Seeded synthetic meshes and PLY I/O (the reference's own assets are not shipped).

S-blob(n): bumpy UV sphere r = 1 + .08 sin7θ cos9φ + .03 sin(23θ+1.3) sin31φ, n x n quads
-> 2 n^2 triangles.  Written as binary little-endian PLY with float x,y,z and
`list uchar int vertex_indices`, the layout procedures/stanfordply_procedure/ply2mesh.cc:32-49 reads.
"""
import numpy as np

BLOB_N = {"70k": 187, "1M": 707, "7M": 1871, "10M": 2236, "buddha": 740}


def blob(n, radius=1.0):
    """Returns (P float32 [V,3], idx int32 [F,3]) of the bumpy sphere with 2*n*n triangles."""
    theta = np.linspace(0.0, np.pi, n + 1)
    phi = np.linspace(0.0, 2 * np.pi, n, endpoint=False)
    T, Ph = np.meshgrid(theta, phi, indexing="ij")
    r = radius * (1 + .08 * np.sin(7 * T) * np.cos(9 * Ph) + .03 * np.sin(23 * T + 1.3) * np.sin(31 * Ph))
    P = np.stack([r * np.sin(T) * np.cos(Ph), r * np.cos(T), r * np.sin(T) * np.sin(Ph)], -1)
    P = P.reshape(-1, 3).astype(np.float32)
    i = np.arange(n)[:, None]
    j = np.arange(n)[None, :]
    a = i * n + j
    b = i * n + (j + 1) % n
    c = (i + 1) * n + j
    d = (i + 1) * n + (j + 1) % n
    idx = np.stack([np.stack([a, c, b], -1), np.stack([b, c, d], -1)], 2).reshape(-1, 3).astype(np.int32)
    return P, idx


def random_tris(ntris, seed=1234):
    """S-random: triangles with centroids uniform in the unit cube, edge ~ N^(-1/3)."""
    rng = np.random.default_rng(seed)
    c = rng.random((ntris, 1, 3))
    e = (rng.random((ntris, 3, 3)) - .5) * (2.0 * ntris ** (-1.0 / 3))
    P = (c + e).reshape(-1, 3).astype(np.float32) - np.float32(.5)
    idx = np.arange(3 * ntris, dtype=np.int32).reshape(-1, 3)
    return P, idx


def quad(size=1.0, y=0.0):
    """2-triangle floor in the xz plane, normal +y."""
    s = size
    P = np.array([[-s, y, -s], [s, y, -s], [s, y, s], [-s, y, s]], np.float32)
    idx = np.array([[0, 2, 1], [0, 3, 2]], np.int32)
    return P, idx


def cube(size=1.0):
    """Flat-shaded cube: 24 vertices (4 per face), 12 triangles, side `size`, centred at the origin —
    the shape of the reference's only shipped mesh (scenes/cube.ply: 24 verts, 6 quads)."""
    h = .5 * size
    P, idx = [], []
    for axis in range(3):
        for sgn in (-1, 1):
            u, v = (axis + 1) % 3, (axis + 2) % 3
            base = len(P)
            for du, dv in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
                p = [0., 0., 0.]
                p[axis], p[u], p[v] = sgn * h, du * h, dv * h
                P.append(p)
            quad_ = [base, base + 1, base + 2, base + 3] if sgn > 0 else [base, base + 3, base + 2, base + 1]
            idx.append([quad_[0], quad_[1], quad_[2]])
            idx.append([quad_[0], quad_[2], quad_[3]])
    return np.asarray(P, np.float32), np.asarray(idx, np.int32)


def write_ply(path, P, idx, uv=None):
    """Binary PLY the reference's StanfordPlyProcedure reads (ply2mesh.cc:32-49); `uv` [V,2] adds the float properties
    uv1 / uv2 it maps to Mesh::SetPointTexture (ply2mesh.cc:42-43,156-161)."""
    P = np.ascontiguousarray(P, dtype="<f4")
    idx = np.ascontiguousarray(idx, dtype="<i4")
    uvp = "property float uv1\nproperty float uv2\n" if uv is not None else ""
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
           "property float z\n%selement face %d\nproperty list uchar int vertex_indices\nend_header\n"
           % (len(P), uvp, len(idx)))
    rec = np.empty(len(idx), dtype=[("n", "u1"), ("v", "<i4", 3)])
    rec["n"] = 3
    rec["v"] = idx
    vert = P if uv is None else np.concatenate([P, np.ascontiguousarray(uv, dtype="<f4")], axis=1)
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(np.ascontiguousarray(vert, dtype="<f4").tobytes())
        f.write(rec.tobytes())


def write_mip(path, img, tilesize=64):
    """Writes the reference's `.mip` texture format (src/fj_mipmap.cc:124-180, 270-320): "MIPM", int32 version 1, width,
    height, nchannels, tilesize, then the row-major tiles of tilesize x tilesize x nchannels float32.  `img` is
    [H, W, C] with H and W multiples of `tilesize`, row 0 = top of the image.  Returns the tile array [ny, nx, ts, ts, C]."""
    img = np.ascontiguousarray(img, np.float32)
    h, w, c = img.shape
    assert h % tilesize == 0 and w % tilesize == 0 and c in (1, 3, 4)
    tiles = img.reshape(h // tilesize, tilesize, w // tilesize, tilesize, c).transpose(0, 2, 1, 3, 4)
    tiles = np.ascontiguousarray(tiles)
    with open(path, "wb") as f:
        f.write(b"MIPM")
        f.write(np.array([1, w, h, c, tilesize], "<i4").tobytes())
        f.write(tiles.astype("<f4").tobytes())
    return tiles


def sphere_uv(P):
    """Longitude / latitude texture coordinates of points around the origin (seam at -x)."""
    P = np.asarray(P, np.float64)
    r = np.linalg.norm(P, axis=1) + 1e-30
    u = np.arctan2(P[:, 2], P[:, 0]) / (2 * np.pi) + .5
    v = np.arccos(np.clip(P[:, 1] / r, -1, 1)) / np.pi
    return np.stack([u, 1 - v], -1).astype(np.float32)


def read_ply(path):
    """Minimal PLY reader (ascii / binary_little_endian; float x,y,z + extra float props;
    polygon faces fan-triangulated like ply2mesh.cc:129-136).  Returns (P float32, idx int32)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").split("\n")
    fmt = None
    elems = []
    for ln in lines:
        w = ln.split()
        if not w:
            continue
        if w[0] == "format":
            fmt = w[1]
        elif w[0] == "element":
            elems.append([w[1], int(w[2]), []])
        elif w[0] == "property":
            elems[-1][2].append(w[1:])
    sz = {"char": 1, "uchar": 1, "short": 2, "ushort": 2, "int": 4, "uint": 4, "float": 4, "double": 8,
          "int8": 1, "uint8": 1, "int16": 2, "uint16": 2, "int32": 4, "uint32": 4, "float32": 4, "float64": 8}
    npt = {"char": "i1", "uchar": "u1", "short": "<i2", "ushort": "<u2", "int": "<i4", "uint": "<u4",
           "float": "<f4", "double": "<f8", "int8": "i1", "uint8": "u1", "int16": "<i2", "uint16": "<u2",
           "int32": "<i4", "uint32": "<u4", "float32": "<f4", "float64": "<f8"}
    P = None
    tris = []
    if fmt == "ascii":
        toks = data[end:].split()
        pos = 0
        for name, cnt, props in elems:
            if name == "vertex":
                nprop = len(props)
                arr = np.array(toks[pos:pos + cnt * nprop], dtype=np.float64).reshape(cnt, nprop)
                names = [p[-1] for p in props]
                P = arr[:, [names.index("x"), names.index("y"), names.index("z")]].astype(np.float32)
                pos += cnt * nprop
            elif name == "face":
                for _ in range(cnt):
                    n = int(toks[pos])
                    v = [int(t) for t in toks[pos + 1:pos + 1 + n]]
                    pos += 1 + n
                    for k in range(n - 2):
                        tris.append((v[0], v[k + 1], v[k + 2]))
    else:
        pos = end
        for name, cnt, props in elems:
            if name == "vertex":
                dt = np.dtype([(p[-1], npt[p[0]]) for p in props])
                arr = np.frombuffer(data, dtype=dt, count=cnt, offset=pos)
                P = np.stack([arr["x"], arr["y"], arr["z"]], -1).astype(np.float32)
                pos += cnt * dt.itemsize
            elif name == "face":
                p = props[0]
                ct, it = p[1], p[2]
                dt3 = np.dtype([("n", npt[ct]), ("v", npt[it], 3)])
                arr = np.frombuffer(data, dtype=dt3, count=cnt, offset=pos)
                if cnt and np.all(arr["n"] == 3):
                    tris = arr["v"].astype(np.int32)
                    pos += cnt * dt3.itemsize
                else:
                    for _ in range(cnt):
                        n = int(np.frombuffer(data, npt[ct], 1, pos)[0])
                        pos += sz[ct]
                        v = np.frombuffer(data, npt[it], n, pos)
                        pos += sz[it] * n
                        for k in range(n - 2):
                            tris.append((int(v[0]), int(v[k + 1]), int(v[k + 2])))
    return P, np.asarray(tris, dtype=np.int32).reshape(-1, 3)
