#!/usr/bin/env python
"""TEST INFRASTRUCTURE: the reference's shipped example scenes (scenes/*.scn, scenes/*.py) as command streams for the
shipped-scene runner (tests/test_shipped_scenes_*.py; the counterpart of the reference's tests/run_all_scenes.py:30-60).

Runs where /root/reference exists (this container; `__graft_entry__.build()` calls it) and writes ONLY into oracle/_ref/shipped/
— git-ignored, so nothing of the reference enters the repo's history, but it travels to the GPU box with the built reference:

  * a `.scn` file is taken as it is; a `.py` script is executed UNCHANGED under Python 3 with this repo's `fujiyama` module
    (fujiyama-renderer_b200/fujiyama.py, the py3 mirror of tools/python_api/fujiyama.py) in print mode, which yields the same
    command stream the reference's shim would pipe into `bin/scene`;
  * asset and plugin paths become placeholders (${PLY}/name.ply, ${MIP}/name.mip, ${PLUGINS}/Name.so, ${OUT}.fb): the assets of
    every shipped scene but the cube live outside the reference tree (INSTALL:56-64) and the runner substitutes seeded synthetic
    stand-ins (SURVEY.md 8c gotcha 3).

In scope: the scenes whose plugins are on the device path (Constant / Plastic / Glass / Pathtracing shaders, StanfordPly and
VelocityGenerator procedures).  Scenes with volumes, curves, point clouds, hair / sss / material shaders are not."""
import contextlib
import io
import os
import re
import runpy
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SCENES = "/root/reference/scenes"
OUT = os.path.join(REPO, "oracle", "_ref", "shipped")

IN_SCOPE = ["happy_buddhas.scn", "xyzrgb_dragon.scn", "teapot.scn",
            "happy_buddhas.py", "xyzrgb_dragon.py", "teapot2.py", "pathtracing.py", "grid_light.py", "dome_light1.py",
            "dome_light2.py", "glassy_happy.py", "transform_motion_blur.py", "camera_motion_blur.py", "bump_mapping.py",
            "mesh_velocity_blur.py", "sphere_light.py"]


def commands_of_script(path):
    """Executes a shipped scenes/*.py unchanged with the py3 `fujiyama` module; returns the command stream it builds."""
    sys.path.insert(0, REPO)
    import __graft_entry__ as entry
    entry.load_package()
    from fujiyama_renderer_b200 import fujiyama as shim
    sys.modules["fujiyama"] = shim
    argv, sys.argv = sys.argv, [path, "-P"]                     # print mode: Run() prints the stream instead of rendering
    buf = io.StringIO()
    cwd = os.getcwd()
    try:
        os.chdir(os.path.dirname(path))
        with contextlib.redirect_stdout(buf):
            runpy.run_path(path, run_name="__main__")
    finally:
        os.chdir(cwd)
        sys.argv = argv
        sys.modules.pop("fujiyama", None)
    return buf.getvalue()


def localize(text):
    out = []
    for line in text.split("\n"):
        t = line.strip()
        if not t or t.startswith("#"):
            continue
        w = t.split()
        if w[0] == "OpenPlugin":
            w[2] = "${PLUGINS}/" + os.path.splitext(os.path.basename(w[2]))[0] + ".so"
        elif w[0] == "SetStringProperty" and w[2] == "filepath":
            w[3] = "${PLY}/" + os.path.basename(w[3])
        elif w[0] == "NewTexture":
            w[2] = "${MIP}/" + os.path.splitext(os.path.basename(w[2]))[0] + ".mip"
        elif w[0] == "SaveFrameBuffer":
            w[2] = "${OUT}.fb"
        out.append(" ".join(w))
    return "\n".join(out) + "\n"


def main():
    if not os.path.isdir(REF_SCENES):
        print("gen_shipped: %s not present, nothing to do" % REF_SCENES)
        return
    os.makedirs(OUT, exist_ok=True)
    for name in IN_SCOPE:
        src = os.path.join(REF_SCENES, name)
        text = open(src).read() if name.endswith(".scn") else commands_of_script(src)
        text = localize(text)
        assert "RenderScene" in text and "${OUT}.fb" in text, name
        with open(os.path.join(OUT, re.sub(r"\.(scn|py)$", lambda m: "_" + m.group(1), name) + ".scn"), "w") as f:
            f.write(text)
    print("gen_shipped: %d scene command streams in %s" % (len(IN_SCOPE), OUT))


if __name__ == "__main__":
    main()
