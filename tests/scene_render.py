"""TEST INFRASTRUCTURE: renders the reference's integration-test scene (reference/model_test.glsl) with a
model's tables on the GPU -- tests/cuda/scene_kernel.cu, a kernel of the TEST tree over the product's
device-side lookups (csrc/kernel_render.cuh; context from Model.render_context()). The scene is not part
of libpas_b200.so."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "cuda", "libpas_test_scene.so")
_lib = None
last_kernel_ms = 0.0


class _SceneView(ctypes.Structure):
    """SceneView of tests/cuda/scene_kernel.cu."""
    _fields_ = [("camera", ctypes.c_double * 3), ("earth_center", ctypes.c_double * 3),
                ("sun_direction", ctypes.c_double * 3), ("sun_size", ctypes.c_double * 2),
                ("sphere_center", ctypes.c_double * 3), ("sphere_radius", ctypes.c_double),
                ("model_from_clip", ctypes.c_double * 9), ("ground_albedo", ctypes.c_double * 3),
                ("sphere_albedo", ctypes.c_double * 3), ("exposure", ctypes.c_double),
                ("use_luminance", ctypes.c_int), ("width", ctypes.c_int), ("height", ctypes.c_int)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "cuda")])
    return LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: build it with make -C tests/cuda (or __graft_entry__.build())")
        lib = ctypes.CDLL(LIB_PATH)
        lib.pas_test_render_scene.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
        _lib = lib
    return _lib


def render_scene(model, view, want_argb: bool = True):
    """(rgb float32 [H, W, 3] before tone mapping, argb uint32 [H, W] or None) for a scene.SceneView."""
    global last_kernel_ms
    if view.width < 1 or view.height < 1 or not (view.sphere_radius > 0.0) or not (view.sun_size[0] > 0.0):
        raise ValueError("bad view")
    ctx = model.render_context(bool(view.use_luminance))     # raises PasError before Init / in the wrong mode
    v = _SceneView()
    for name in ("camera", "earth_center", "sun_direction", "sun_size", "sphere_center", "model_from_clip",
                 "ground_albedo", "sphere_albedo"):
        getattr(v, name)[:] = list(getattr(view, name))
    v.sphere_radius, v.exposure = view.sphere_radius, view.exposure
    v.use_luminance, v.width, v.height = int(view.use_luminance), view.width, view.height
    rgb = np.empty((view.height, view.width, 3), np.float32)
    argb = np.empty((view.height, view.width), np.uint32) if want_argb else None
    ms = ctypes.c_float(0)
    if model.device is not None:
        import torch
        torch.cuda.set_device(model.device)
    rc = _load().pas_test_render_scene(ctx, len(ctx), ctypes.byref(v), ctypes.sizeof(v), rgb.ctypes.data,
                                       argb.ctypes.data if want_argb else None, ctypes.byref(ms))
    if rc != 0:
        raise RuntimeError(f"scene kernel failed: CUDA error {rc}")
    last_kernel_ms = ms.value
    return rgb, argb
