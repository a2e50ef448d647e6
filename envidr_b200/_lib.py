"""ctypes binding of libenvidr_b200.so (the C ABI declared in include/envidr_b200.h).

The product path has no fallback: if the shared library is missing or a call fails, an exception is
raised.  Function prototypes are parsed from the header so that the binding cannot drift from it.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "envidr_b200.h")
LIB_PATH = os.path.join(_HERE, "_lib", "libenvidr_b200.so")

ENVIDR_MAX_LAYERS = 8


class MlpLayer(ctypes.Structure):
    _fields_ = [("weight", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("in_dim", ctypes.c_uint32), ("out_dim", ctypes.c_uint32)]


class Field(ctypes.Structure):
    _fields_ = [
        ("embeddings", ctypes.c_void_p), ("offsets", ctypes.c_void_p),
        ("num_levels", ctypes.c_uint32), ("level_dim", ctypes.c_uint32), ("base_resolution", ctypes.c_uint32),
        ("log2_per_level_scale", ctypes.c_float), ("bound", ctypes.c_float), ("enabled_levels", ctypes.c_int32),
        ("n_sdf", ctypes.c_uint32), ("n_env", ctypes.c_uint32), ("n_diffuse", ctypes.c_uint32), ("n_color", ctypes.c_uint32),
        ("n_renv", ctypes.c_uint32),
        ("sdf", MlpLayer * ENVIDR_MAX_LAYERS), ("env", MlpLayer * ENVIDR_MAX_LAYERS), ("diffuse", MlpLayer * ENVIDR_MAX_LAYERS),
        ("color", MlpLayer * ENVIDR_MAX_LAYERS), ("renv", MlpLayer * ENVIDR_MAX_LAYERS),
        ("geo_feat_dim", ctypes.c_uint32), ("ide_degree", ctypes.c_uint32),
        ("beta", ctypes.c_float), ("density_scale", ctypes.c_float),
        ("roughness_bias", ctypes.c_float), ("roughness_act_scale", ctypes.c_float), ("roughness_scale", ctypes.c_float),
        ("diffuse_kappa_inv", ctypes.c_float), ("light_intensity_scale", ctypes.c_float), ("intensity_scale", ctypes.c_float),
        ("indir_roughness_thresh", ctypes.c_float), ("learn_indir_blend", ctypes.c_int32),
        ("has_env_rot", ctypes.c_int32), ("env_rot", ctypes.c_float * 9),
        ("packed", ctypes.c_void_p), ("packed_bytes", ctypes.c_uint64),
        ("precision", ctypes.c_int32), ("rec_unrotated", ctypes.c_int32),
        ("scratch", ctypes.c_void_p), ("scratch_samples", ctypes.c_uint64),
    ]


class NeusLayer(ctypes.Structure):
    _fields_ = [("img", ctypes.c_void_p), ("imgT", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("in_dim", ctypes.c_uint32), ("out_dim", ctypes.c_uint32)]


class NeusNet(ctypes.Structure):
    _fields_ = [("layers", NeusLayer * 8), ("head_row", ctypes.c_void_p), ("n_layers", ctypes.c_uint32), ("skip_layer", ctypes.c_int32),
                ("multires", ctypes.c_uint32), ("beta", ctypes.c_float)]


class FieldOut(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("sigma", "rgb", "normal", "sdf", "c_diffuse", "c_specular", "roughness", "grad_x")]


class RenderOpts(ctypes.Structure):
    _fields_ = [("bound", ctypes.c_float), ("dt_gamma", ctypes.c_float), ("T_thresh", ctypes.c_float), ("min_near", ctypes.c_float),
                ("max_steps", ctypes.c_uint32), ("cascade", ctypes.c_uint32), ("grid_size", ctypes.c_uint32),
                ("aabb", ctypes.c_float * 6), ("bg_color", ctypes.c_float * 3),
                ("geometry_only", ctypes.c_int32), ("input_alpha", ctypes.c_int32), ("n_step_floor", ctypes.c_uint32),
                ("n_step_cap", ctypes.c_uint32)]


class SampleLog(ctypes.Structure):
    _fields_ = [("rec", ctypes.c_void_p), ("sigma", ctypes.c_void_p), ("delta", ctypes.c_void_p), ("ray", ctypes.c_void_p),
                ("seq", ctypes.c_void_p), ("capacity", ctypes.c_uint64)]


class RenderOut(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("image", "depth", "weights_sum", "normal_image", "diffuse_image", "specular_image",
                                              "roughness_image", "sample_count", "log")]


_SCALARS = {"uint32_t": ctypes.c_uint32, "uint64_t": ctypes.c_uint64, "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "int": ctypes.c_int,
            "float": ctypes.c_float, "double": ctypes.c_double, "envidr_stream_t": ctypes.c_void_p}
_RET = {"int": ctypes.c_int, "uint64_t": ctypes.c_uint64, "const char*": ctypes.c_char_p}


def parse_header(path: str = HEADER):
    """Return {function name: (restype, [argtypes])} for every prototype declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?m)^(const char\*|int|uint64_t)\s+(envidr_\w+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a or "[" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_SCALARS[ty])
        protos[name] = (_RET[ret], argtypes)
    return protos


class EnvidrError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the CUDA library (building is the job of __graft_entry__.build / python -m envidr_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EnvidrError(f"{LIB_PATH} is missing: build it with `python -m envidr_b200.build` (no CPU fallback exists)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (ret, argtypes) in parse_header().items():
            fn = getattr(l, name)          # AttributeError here means the header and the library disagree
            fn.restype, fn.argtypes = ret, argtypes
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().envidr_last_error().decode()
        raise EnvidrError(f"{what or 'envidr'} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
