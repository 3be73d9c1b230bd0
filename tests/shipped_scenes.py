"""Runner support for the reference's shipped example scenes (the counterpart of the reference's tests/run_all_scenes.py:30-60):
the command streams oracle/gen_shipped.py wrote into oracle/_ref/shipped/ (the shipped `.scn` files, and the shipped `.py`
scripts executed unchanged on this repo's py3 `fujiyama` module), seeded synthetic stand-ins for the assets the reference keeps
outside its tree (INSTALL:56-64), and the substitution that points a stream at them."""
import glob
import os
import zlib

import numpy as np

import scenekit as sk

SHIPPED_DIR = os.path.join(sk.REPO, "oracle", "_ref", "shipped")

# bumpy-sphere stand-ins for the scanned models (radius, blob resolution), resting on the floor at y = 0
MODELS = {"happy": (.75, 28), "bunny": (.7, 26), "armadillo": (.8, 24), "horse": (.7, 22), "dragon": (.9, 30),
          "xyzrgb_dragon": (.9, 32), "teapot": (.8, 20)}


def streams():
    return sorted(glob.glob(os.path.join(SHIPPED_DIR, "*.scn")))


def make_assets(root):
    """Seeded synthetic PLY / mip stand-ins, by the file names the shipped scenes ask for."""
    sk.pkg()
    from fujiyama_renderer_b200 import synth
    ply, mip = os.path.join(root, "ply"), os.path.join(root, "mip")
    os.makedirs(ply, exist_ok=True)
    os.makedirs(mip, exist_ok=True)
    for name, (radius, n) in MODELS.items():
        P, idx = synth.blob(n, radius)
        P = P.copy(); P[:, 1] += np.float32(radius * 1.12)
        synth.write_ply(os.path.join(ply, name + ".ply"), P, idx)
    P, idx = synth.quad(40.0, 0.0)
    synth.write_ply(os.path.join(ply, "floor.ply"), P, idx)
    P, idx = synth.blob(24, 300.0)                       # the environment dome the scenes map a texture on
    synth.write_ply(os.path.join(ply, "dome.ply"), P, idx, uv=synth.sphere_uv(P))
    P, idx = synth.blob(20, 1.0)
    synth.write_ply(os.path.join(ply, "sphere.ply"), P, idx)
    synth.write_ply(os.path.join(ply, "sphere_uv.ply"), P, idx, uv=synth.sphere_uv(P))
    for path in streams():
        for line in open(path):
            w = line.split()
            if w and w[0] == "NewTexture":
                name = os.path.basename(w[2])
                dst = os.path.join(mip, name)
                if not os.path.exists(dst):
                    rng = np.random.default_rng(zlib.crc32(name.encode()))
                    yy, xx = np.mgrid[0:128, 0:256]
                    base = 0.35 + 0.3 * np.sin(xx / 256.0 * 2 * np.pi * rng.integers(1, 4))[..., None] * np.ones(3)
                    sky = np.clip(1.2 - yy / 128.0, 0.1, 1.5)[..., None] * rng.uniform(0.6, 1.0, 3)
                    img = (base * 0.4 + sky * 0.6 + 0.05 * rng.random((128, 256, 3))).astype(np.float32)
                    synth.write_mip(dst, img)
    return {"PLY": ply, "MIP": mip}


def prepare(text, assets, plugin_dir, out_base, res=(160, 120), spp=(2, 2), threads=None):
    """Points a stream at the stand-ins and the given plugin directory, renders at `res` with `spp` pixel samples (the
    reference's own runner shrinks the resolution the same way, tests/run_all_scenes.py:36,50)."""
    out = []
    for line in text.split("\n"):
        w = line.split()
        if len(w) >= 5 and w[0] == "SetProperty2" and w[2] == "resolution":
            w[3], w[4] = str(res[0]), str(res[1])
        if len(w) >= 5 and w[0] == "SetProperty2" and w[2] == "pixelsamples":
            w[3], w[4] = str(spp[0]), str(spp[1])
        if w and w[0] == "RenderScene" and threads:
            out.append("SetProperty1 %s use_max_thread 0" % w[1])
            out.append("SetProperty1 %s thread_count %d" % (w[1], threads))
        out.append(" ".join(w))
    t = "\n".join(out) + "\n"
    return (t.replace("${PLUGINS}", plugin_dir).replace("${PLY}", assets["PLY"]).replace("${MIP}", assets["MIP"])
            .replace("${OUT}", out_base))


def is_stochastic(text):
    return any(k in text for k in ("GridLight", "SphereLight", "PathtracingShader"))
