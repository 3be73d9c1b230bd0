"""Thin Python handle on one libfjgpu context (one per GPU / per rank).

This is plumbing over the C-ABI of include/fjgpu.h: it owns an `fjgpu_context*`, forwards flat
scene arrays and returns numpy frames.  Every call raises `FjGpuError` with `fjgpu_last_error` on a
non-zero status; there is no CPU fallback (fjgpu_create fails without a B200).
"""
import ctypes as C

import numpy as np

from . import abi


class FjGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fjgpu error %d: %s" % (code, msg))
        self.code = code


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def tile_array(tiles):
    arr = (abi.Tile * max(1, len(tiles)))()
    for i, t in enumerate(tiles):
        arr[i].id, arr[i].xmin, arr[i].ymin, arr[i].xmax, arr[i].ymax = t
    return arr


class Device:
    def __init__(self, ordinal=0):
        self.lib = abi.load_fjgpu()
        self.ctx = C.c_void_p()
        rc = self.lib.fjgpu_create(ordinal, C.byref(self.ctx))
        if rc != 0:
            raise FjGpuError(rc, (self.lib.fjgpu_last_error(None) or b"").decode())

    def close(self):
        if self.ctx:
            self.lib.fjgpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise FjGpuError(rc, (self.lib.fjgpu_last_error(self.ctx) or b"").decode())

    # ---- scene
    def mesh(self, mesh_id, P, N, idx, group=None, velocity=None):
        P = np.ascontiguousarray(P, np.float64)
        idx = np.ascontiguousarray(idx, np.int32).reshape(-1)
        Np = None
        if N is not None:
            N = np.ascontiguousarray(N, np.float64)
            Np = _dp(N)
        gp = None
        if group is not None:
            group = np.ascontiguousarray(group, np.int32)
            gp = _ip(group)
        if velocity is not None:      # Mesh::velocity_: moving triangles (fjgpu_mesh_upload_velocity)
            velocity = np.ascontiguousarray(velocity, np.float64)
            assert velocity.shape == P.shape
            self._ck(self.lib.fjgpu_mesh_upload_velocity(self.ctx, mesh_id, _dp(P), Np, len(P), _ip(idx), gp, len(idx) // 3, _dp(velocity)))
            return
        self._ck(self.lib.fjgpu_mesh_upload(self.ctx, mesh_id, _dp(P), Np, len(P), _ip(idx), gp, len(idx) // 3))

    def load_structs(self, st):
        """`st`: the flat struct dict (meshes, instances, group_offsets/ids, shaders, lights, camera)."""
        vel = st.get("mesh_velocity", {})
        for mid in st.get("velocity_meshes", []):
            if mid not in vel:
                raise FjGpuError(-1, "mesh %d is marked as moving but carries no velocities (st['mesh_velocity'])" % mid)
        for mid, P, N, idx in st["meshes"]:
            self.mesh(mid, P, N, idx, velocity=vel.get(mid))
        t0, t1 = st.get("time_range", (0.0, 1.0))
        self._ck(self.lib.fjgpu_shutter_set(self.ctx, float(t0), float(t1)))
        for mid, uv in st.get("mesh_uv", {}).items():
            uv = np.ascontiguousarray(uv, np.float32)
            self._ck(self.lib.fjgpu_mesh_set_uv(self.ctx, mid, uv.ctypes.data_as(C.POINTER(C.c_float)), len(uv)))
        if "textures" in st:
            self._ck(self.lib.fjgpu_textures_set(self.ctx, st["ntextures"], st["textures"]))
        self._ck(self.lib.fjgpu_shaders_set(self.ctx, st["nshaders"], st["shaders"]))
        off = np.ascontiguousarray(st["group_offsets"], np.int32)
        ids = np.ascontiguousarray(st["group_ids"], np.int32)
        self._ck(self.lib.fjgpu_groups_set(self.ctx, len(off) - 1, _ip(off), _ip(ids)))
        self._ck(self.lib.fjgpu_instances_set(self.ctx, st["ninstances"], st["instances"]))
        self._ck(self.lib.fjgpu_lights_set(self.ctx, st["nlights"], st["lights"]))
        self._ck(self.lib.fjgpu_camera_set(self.ctx, C.byref(st["camera"])))
        # motion blur: matrices at every entry of the frame's time table (st["motion"]: index -> (fwd [n,16], inv [n,16]),
        # index -1 = the camera); a scene without them resets what an earlier scene left in the context
        motion = st.get("motion", {})
        for i, (fwd, inv) in motion.items():
            fwd, inv = np.ascontiguousarray(fwd, np.float64), np.ascontiguousarray(inv, np.float64)
            if i < 0:
                self._ck(self.lib.fjgpu_camera_motion_set(self.ctx, len(fwd), _dp(fwd)))
            else:
                self._ck(self.lib.fjgpu_instance_motion_set(self.ctx, i, len(fwd), _dp(fwd), _dp(inv)))
        if -1 not in motion:
            self._ck(self.lib.fjgpu_camera_motion_set(self.ctx, 0, None))
        for i in range(st["ninstances"]):
            if i not in motion:
                self._ck(self.lib.fjgpu_instance_motion_set(self.ctx, i, 0, None, None))

    def info(self):
        i = abi.SceneInfo()
        self._ck(self.lib.fjgpu_scene_info_get(self.ctx, C.byref(i)))
        return i

    # ---- frame
    def render(self, params, tiles, frame=None):
        """fjgpu_render_tiles: host frame [yres, xres, 4] float32 (created zeroed when None)."""
        ta = tiles if not isinstance(tiles, (list, tuple)) else tile_array(tiles)
        n = len(tiles)
        if frame is None:
            frame = np.zeros((params.yres, params.xres, 4), np.float32)
        stats = abi.Stats()
        self._ck(self.lib.fjgpu_render_tiles(self.ctx, C.byref(params), ta, n, _fp(frame), C.byref(stats)))
        return frame, stats

    def render_resident(self, params, tiles, want_stats=True):
        ta = tiles if not isinstance(tiles, (list, tuple)) else tile_array(tiles)
        stats = abi.Stats()
        self._ck(self.lib.fjgpu_render_tiles_resident(self.ctx, C.byref(params), ta, len(tiles),
                                                      C.byref(stats) if want_stats else None))
        return stats

    def render_to_device_blocks(self, params, tiles, tile_w, tile_h, dev_ptr):
        ta = tiles if not isinstance(tiles, (list, tuple)) else tile_array(tiles)
        stats = abi.Stats()
        self._ck(self.lib.fjgpu_render_tiles_device(self.ctx, C.byref(params), ta, len(tiles), tile_w, tile_h,
                                                    C.c_void_p(dev_ptr), C.byref(stats)))
        return stats

    def assemble_frame(self, gathered_ptr, nranks, tile_w, tile_h, tiles, xres, yres, frame=None):
        """fjgpu_assemble_frame: all-gathered tile blocks (device pointer, rank order) -> host frame [yres, xres, 4]."""
        ta = tiles if not isinstance(tiles, (list, tuple)) else tile_array(tiles)
        if frame is None:
            frame = np.zeros((yres, xres, 4), np.float32)
        self._ck(self.lib.fjgpu_assemble_frame(self.ctx, C.c_void_p(gathered_ptr), nranks, tile_w, tile_h, ta, len(tiles),
                                               xres, yres, frame.ctypes.data_as(C.c_void_p)))
        return frame

    def trace_closest(self, group, orig, dirs, tmin, tmax, flags=0):
        orig = np.ascontiguousarray(orig, np.float64)
        dirs = np.ascontiguousarray(dirs, np.float64)
        n = len(orig)
        tmin = np.ascontiguousarray(np.broadcast_to(tmin, (n,)), np.float64)
        tmax = np.ascontiguousarray(np.broadcast_to(tmax, (n,)), np.float64)
        t, u, v = np.zeros(n), np.zeros(n), np.zeros(n)
        prim, inst = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._ck(self.lib.fjgpu_trace_closest(self.ctx, group, n, _dp(orig), _dp(dirs), _dp(tmin), _dp(tmax), flags,
                                              _dp(t), _dp(u), _dp(v), _ip(prim), _ip(inst)))
        return t, u, v, prim, inst

    def tile_samples(self, params, tile):
        t = abi.Tile(*tile)
        n = C.c_int32(0)
        self._ck(self.lib.fjgpu_render_tile_samples(self.ctx, C.byref(params), C.byref(t), 0, None, None, C.byref(n)))
        uv = np.zeros((n.value, 2))
        rgba = np.zeros((n.value, 4), np.float32)
        self._ck(self.lib.fjgpu_render_tile_samples(self.ctx, C.byref(params), C.byref(t), n.value, _dp(uv), _fp(rgba),
                                                    C.byref(n)))
        return uv, rgba


def render_frame_multi(devices, params, tiles, frame=None):
    """fjgpu_render_frame_multi: ONE process, one Device per GPU with the same scene loaded on each; tiles dealt round-robin,
    one NCCL all-gather, rank 0 assembles.  Returns (frame [yres, xres, 4], [Stats per device])."""
    lib = devices[0].lib
    ta = tiles if not isinstance(tiles, (list, tuple)) else tile_array(tiles)
    if frame is None:
        frame = np.zeros((params.yres, params.xres, 4), np.float32)
    ctxs = (C.c_void_p * len(devices))(*[d.ctx for d in devices])
    stats = (abi.Stats * len(devices))()
    rc = lib.fjgpu_render_frame_multi(ctxs, len(devices), C.byref(params), ta, len(tiles), _fp(frame), stats)
    if rc != 0:
        raise FjGpuError(rc, (lib.fjgpu_last_error(devices[0].ctx) or b"").decode())
    return frame, list(stats)
