"""Python-3 `fujiyama` module: the `SceneInterface` surface of the reference's Python-2 shim
(tools/python_api/fujiyama.py:15-361) over libfjscene.so, so that scene scripts written as

    import fujiyama
    si = fujiyama.SceneInterface()
    si.OpenPlugin('PlasticShader', '...'); si.NewMesh('mesh1'); ... ; si.RenderScene('ren1'); si.Run()

run here unchanged.  Every method appends one line of the `.scn` command grammar
(tools/scene_parser/command.cc:502-543); `Run()` executes the accumulated text in-process through
`fjscene_parse_text` (the reference pipes it to the `scene` binary, fujiyama.py:97-98), which renders
on the GPU through the C-ABI of include/fjgpu.h.  There is no CPU renderer behind this module.
"""
import argparse
import ctypes as C
import os

import numpy as np

from . import abi

# command -> number of arguments after the command name (the shim's one-line formatters, fujiyama.py:136-333)
_COMMANDS = {
    "RenderScene": 1, "RunProcedure": 1, "SaveFrameBuffer": 2, "AddObjectToGroup": 2,
    "NewObjectInstance": 2, "NewFrameBuffer": 2, "NewObjectGroup": 1, "NewPointCloud": 1, "NewTurbulence": 1,
    "NewProcedure": 2, "NewRenderer": 1, "NewTexture": 2, "NewCamera": 2, "NewShader": 2, "NewVolume": 1,
    "NewCurve": 1, "NewLight": 2, "NewMesh": 1,
    "AssignShader": 3, "AssignTexture": 3, "AssignCamera": 2, "AssignObjectGroup": 3, "AssignPointCloud": 3,
    "AssignFrameBuffer": 2, "AssignTurbulence": 3, "AssignVolume": 3, "AssignCurve": 3, "AssignMesh": 3,
    "SetProperty1": 3, "SetProperty2": 4, "SetProperty3": 5, "SetProperty4": 6, "SetStringProperty": 3,
    "SetSampleProperty3": 6, "ShowPropertyList": 1,
}

_lib = None


def load_fjscene():
    """Loads libfjscene.so (and libfjgpu.so before it).  Raises if either is not built."""
    global _lib
    if _lib is not None:
        return _lib
    abi.load_fjgpu()
    if not os.path.exists(abi.LIBFJSCENE):
        raise RuntimeError("libfjscene.so is not built (%s): run __graft_entry__.build()" % abi.LIBFJSCENE)
    lib = C.CDLL(abi.LIBFJSCENE, mode=C.RTLD_GLOBAL)
    vp, i32p = C.c_void_p, C.POINTER(C.c_int32)
    lib.fjscene_parser_new.restype = vp
    lib.fjscene_parser_free.argtypes = [vp]
    lib.fjscene_parse_line.argtypes = [vp, C.c_char_p]
    lib.fjscene_parse_text.argtypes = [vp, C.c_char_p]
    lib.fjscene_parse_file.argtypes = [vp, C.c_char_p]
    lib.fjscene_lookup.argtypes = [vp, C.c_char_p]
    lib.fjscene_lookup.restype = C.c_long
    lib.fjscene_set_echo.argtypes = [vp, C.c_int]
    lib.fjscene_mesh_set.argtypes = [C.c_long, C.POINTER(C.c_double), C.c_int32, i32p, C.c_int32]
    lib.fjscene_framebuffer.argtypes = [C.c_long, i32p, i32p, i32p]
    lib.fjscene_framebuffer.restype = C.POINTER(C.c_float)
    lib.fjscene_last_stats.argtypes = [C.POINTER(abi.Stats), C.POINTER(abi.SceneInfo), C.POINTER(C.c_double)]
    lib.fjscene_set_device.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.fjscene_set_resident.argtypes = [C.c_int]
    lib.fjscene_set_resend.argtypes = [C.c_int]
    lib.fjscene_set_gpu_count.argtypes = [C.c_int]
    lib.fjscene_assemble_gathered.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.fjscene_set_device_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.fjscene_last_resend_bytes.restype = C.c_uint64
    lib.fjscene_instance_matrices.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.fjscene_mesh_normals.argtypes = [C.c_long, C.POINTER(C.c_double), C.c_int32]
    lib.fjscene_mesh_velocity.argtypes = [C.c_long, C.POINTER(C.c_double), C.c_int32]
    lib.fjscene_last_message.restype = C.c_char_p
    i32p_ = C.POINTER(C.c_int32)
    lib.fjscene_flatten.argtypes = [C.c_long, i32p_, i32p_, i32p_, i32p_]
    lib.fjscene_flat_instance.argtypes = [C.c_int32, C.POINTER(abi.Instance)]
    lib.fjscene_flat_light.argtypes = [C.c_int32, C.POINTER(abi.Light)]
    lib.fjscene_flat_shader.argtypes = [C.c_int32, C.POINTER(abi.Shader)]
    lib.fjscene_flat_tile.argtypes = [C.c_int32, C.POINTER(abi.Tile)]
    lib.fjscene_flat_frame.argtypes = [C.POINTER(abi.Camera), C.POINTER(abi.RenderParams)]
    lib.fjscene_lerp_transform.argtypes = [C.c_long, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.fjscene_make_transform.argtypes = [C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 5
    _lib = lib
    return lib


class SceneError(RuntimeError):
    pass


class Session:
    """One open scene in libfjscene: feed `.scn` text incrementally, render, read frames back."""

    def __init__(self, echo=False, device=0, rank=0, world_size=1):
        self.lib = load_fjscene()
        self.lib.fjscene_set_device(device, rank, world_size)
        self.lib.fjscene_set_resident(0)
        self.lib.fjscene_set_resend(0)
        self.lib.fjscene_set_gpu_count(0)
        self.lib.fjscene_set_device_blocks(None, 0, 0)
        self.p = self.lib.fjscene_parser_new()
        self.lib.fjscene_set_echo(self.p, 1 if echo else 0)

    def close(self):
        if self.p:
            self.lib.fjscene_parser_free(self.p)
            self.p = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def run(self, text):
        if self.lib.fjscene_parse_text(self.p, text.encode()) != 0:
            raise SceneError((self.lib.fjscene_last_message() or b"scene command failed").decode())

    def id(self, name):
        v = self.lib.fjscene_lookup(self.p, name.encode())
        if v < 0:
            raise KeyError(name)
        return v

    def set_mesh(self, name, P, idx):
        """Fills mesh `name` directly (what a Procedure plugin does through Mesh::Set*)."""
        P = np.ascontiguousarray(P, np.float64)
        idx = np.ascontiguousarray(idx, np.int32).reshape(-1)
        if self.lib.fjscene_mesh_set(self.id(name), P.ctypes.data_as(C.POINTER(C.c_double)), len(P),
                                     idx.ctypes.data_as(C.POINTER(C.c_int32)), len(idx) // 3) != 0:
            raise SceneError("bad mesh arrays")

    def set_resident(self, on):
        self.lib.fjscene_set_resident(1 if on else 0)

    def set_resend(self, on):
        self.lib.fjscene_set_resend(1 if on else 0)

    def set_device_blocks(self, dev_ptr, tile_w, tile_h):
        self.lib.fjscene_set_device_blocks(C.c_void_p(dev_ptr) if dev_ptr else None, tile_w, tile_h)

    def set_gpu_count(self, n):
        """GPUs this ONE process renders a frame on (fjgpu_render_frame_multi); 0 = FJ_GPU_COUNT or 1."""
        self.lib.fjscene_set_gpu_count(int(n))

    def assemble_gathered(self, gathered_ptr, nranks, tile_w, tile_h, host_ptr):
        """Rank 0 after the all-gather: blocks -> frame on the device, one copy to `host_ptr` (pinned memory is written directly)."""
        if self.lib.fjscene_assemble_gathered(C.c_void_p(gathered_ptr), nranks, tile_w, tile_h, C.c_void_p(host_ptr)) != 0:
            raise SceneError((self.lib.fjscene_last_message() or b"assemble failed").decode())

    def resend_bytes(self):
        return int(self.lib.fjscene_last_resend_bytes())

    def framebuffer(self, name):
        w, h, c = C.c_int32(), C.c_int32(), C.c_int32()
        ptr = self.lib.fjscene_framebuffer(self.id(name), C.byref(w), C.byref(h), C.byref(c))
        if not ptr or w.value == 0:
            raise SceneError("framebuffer %s is empty" % name)
        return np.ctypeslib.as_array(ptr, shape=(h.value, w.value, c.value)).copy()

    def stats(self):
        s, i, up = abi.Stats(), abi.SceneInfo(), C.c_double()
        self.lib.fjscene_last_stats(C.byref(s), C.byref(i), C.byref(up))
        return s, i, up.value


class SceneInterface:
    """Drop-in for tools/python_api/fujiyama.py's class of the same name (Python 3)."""

    def __init__(self, argv=None):
        self.commands = []
        ap = argparse.ArgumentParser()
        ap.add_argument("-P", "--print", dest="p", action="store_true",
                        help="force to print scene descriptions instead of running")
        ap.add_argument("-R", "--resolution", dest="res", nargs=2, help="override resolution")
        ap.add_argument("-S", "--pixelsamples", dest="samples", nargs=2, help="override pixel samples")
        self.args, _ = ap.parse_known_args(argv)
        self.session = None

    def __getattr__(self, name):
        n = _COMMANDS.get(name)
        if n is None:
            raise AttributeError(name)

        def emit(*args):
            if len(args) != n:
                raise TypeError("%s takes %d arguments" % (name, n))
            self.commands.append(" ".join([name] + [str(a) for a in args]))
        return emit

    def Comment(self, comment):
        self.commands.append("# %.128s" % comment)

    def OpenPlugin(self, name, plugin_path):
        root, ext = os.path.splitext(plugin_path)
        path = plugin_path if ext == ".so" else plugin_path + ".so"          # fujiyama.py:143-158
        self.commands.append("OpenPlugin %s %s" % (name, path))

    def RenderScene(self, renderer):
        if self.args.res:                                                     # fujiyama.py:160-170
            self.SetProperty2(renderer, "resolution", self.args.res[0], self.args.res[1])
        if self.args.samples:
            self.SetProperty2(renderer, "pixelsamples", self.args.samples[0], self.args.samples[1])
        self.commands.append("RenderScene %s" % renderer)

    def Print(self):
        for c in self.commands:
            print(c)

    def Run(self, echo=True):
        if self.args.p:
            self.Print()
            return
        self.session = Session(echo=echo)
        self.session.run("\n".join(self.commands) + "\n")
