"""fujiyama-renderer_b200 — B200-native hot path of the Fujiyama renderer.

csrc/  hand-written sm_100a CUDA kernels + the extern "C" ABI of include/fjgpu.h (libfjgpu.so)
host/  C++ host mirror of the reference's fj_scene_interface (`Si*`) + `.scn` parser (libfjscene.so, fjscene)
*.py   ctypes bindings, synthetic scene generators, .fb I/O, tile sharding for multi-GPU

The directory name has a hyphen (the repo contract); import it with
`tests/scenekit.pkg()` / `__graft_entry__.load_package()` as `fujiyama_renderer_b200`.
"""
__all__ = ["abi", "synth", "fbio"]
