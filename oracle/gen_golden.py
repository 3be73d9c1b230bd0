#!/usr/bin/env python
"""TEST INFRASTRUCTURE: regenerates tests/golden/ from the UNMODIFIED reference built in oracle/_ref
(make -f oracle/Makefile.ref).  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py

 * tests/golden/ref_vectors.json  per-function vectors printed by oracle/ref_probe.cc linked against the
                                  reference's libscene.so (XorShift, TriRayIntersect, BoxRayIntersect,
                                  make_transform_matrix/MatInverse, Camera::GetRay, FixedGridSampler,
                                  Gaussian filter, SlFresnel/SlReflect/SlRefract, Mesh::ComputeNormals)
 * tests/golden/ref_images.npz    whole frames rendered by the reference binary, thread_count 1
                                  (float32 arrays parsed from its .fb text output), for tests/golden_scenes.py
 * tests/golden/config2_region.npz  (`python oracle/gen_golden.py config2` makes only this one) the centre 22x4 tiles (both silhouettes) of
                                  BASELINE config 2 at FULL size (69 938 triangles, plastic, 1280x720, 16 spp) rendered by the
                                  reference with render_region: what tests/test_fullsize_gpu.py holds the device frame to
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
import scenekit as sk  # noqa: E402
import golden_scenes  # noqa: E402


def config2_region(gold):
    import workloads
    sk.pkg()
    from fujiyama_renderer_b200 import scenes
    desc = workloads.plastic_blob()
    region = scenes.center_region(desc.ren["resolution"], 32, 22, 4)
    with tempfile.TemporaryDirectory() as tmp:
        img, secs = sk.reference_render(desc, tmp, threads=min(8, os.cpu_count() or 1), region=region)
    x0, y0, x1, y1 = region
    crop = np.ascontiguousarray(img[y0:y1, x0:x1])
    assert not img[:y0].any() and not img[y1:].any()          # the reference wrote only the region
    np.savez_compressed(os.path.join(gold, "config2_region.npz"), region=np.asarray(region, np.int32), image=crop)
    print("config2 region %s  %.1fs  mean %.6f  alpha %.3f" % (list(region), secs, crop.mean(), crop[..., 3].mean()))


def main():
    subprocess.check_call(["make", "-f", "oracle/Makefile.ref", "-j8"], cwd=REPO, stdout=subprocess.DEVNULL)
    if sys.argv[1:] == ["config2"]:
        config2_region(os.path.join(REPO, "tests", "golden"))
        return
    if len(sys.argv) == 3 and sys.argv[1] == "add":        # one more scene into ref_images.npz, the others untouched
        path = os.path.join(REPO, "tests", "golden", "ref_images.npz")
        imgs = dict(np.load(path))
        with tempfile.TemporaryDirectory() as tmp:
            imgs[sys.argv[2]], secs = sk.reference_render(golden_scenes.SCENES[sys.argv[2]](), os.path.join(tmp, sys.argv[2]), threads=1)
        np.savez_compressed(path, **imgs)
        print("%-16s %s  %.3fs  mean %.6f" % (sys.argv[2], imgs[sys.argv[2]].shape, secs, imgs[sys.argv[2]].mean()))
        return
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(sk.REF_DIR, "lib"))
    txt = subprocess.check_output([os.path.join(sk.REF_DIR, "bin", "ref_probe"), "vectors"], env=env)
    gold = os.path.join(REPO, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    with open(os.path.join(gold, "ref_vectors.json"), "wb") as f:
        f.write(txt)
    imgs = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, make in list(golden_scenes.SCENES.items()) + list(golden_scenes.ORACLE_ONLY.items()):
            img, secs = sk.reference_render(make(), os.path.join(tmp, name), threads=1)
            imgs[name] = img
            print("%-16s %s  %.3fs  mean %.6f" % (name, img.shape, secs, img.mean()))
    np.savez_compressed(os.path.join(gold, "ref_images.npz"), **imgs)
    config2_region(gold)


if __name__ == "__main__":
    main()
