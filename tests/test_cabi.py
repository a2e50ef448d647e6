"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/envidr_b200.h
declares, the ctypes struct mirrors have the C sizes, and host-only entry points behave (no GPU compute here)."""
import ctypes
import os

import numpy as np
import pytest

from envidr_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def test_header_symbols_all_exported(lib):
    protos = _lib.parse_header()
    expected = {"envidr_near_far_from_aabb", "envidr_sph_from_ray", "envidr_morton3D", "envidr_morton3D_invert", "envidr_packbits",
                "envidr_get_scatter_idx", "envidr_march_rays_train", "envidr_composite_rays_train_forward",
                "envidr_composite_rays_train_backward", "envidr_march_rays", "envidr_composite_rays", "envidr_hash_encode_forward",
                "envidr_hash_encode_backward", "envidr_hash_encode_second_backward", "envidr_grid_encode_forward",
                "envidr_grid_encode_backward", "envidr_freq_encode_forward", "envidr_freq_encode_backward", "envidr_sh_encode_forward",
                "envidr_sh_encode_backward", "envidr_ide_encode_forward", "envidr_field_forward", "envidr_field_pack", "envidr_render_rays"}
    assert expected <= set(protos), expected - set(protos)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert lib.envidr_version() == 103
    assert {"envidr_density_grid_update", "envidr_mark_untrained_grid", "envidr_adam_step", "envidr_get_rays", "envidr_train_loss_forward",
            "envidr_train_loss_backward", "envidr_pow2_scales"} <= set(protos)


def test_struct_mirrors_match_c_layout(lib):
    sizes = (ctypes.c_uint32 * 5)()
    assert lib.envidr_abi_sizes(sizes) == 0
    assert list(sizes) == [ctypes.sizeof(_lib.MlpLayer), ctypes.sizeof(_lib.Field), ctypes.sizeof(_lib.FieldOut),
                           ctypes.sizeof(_lib.RenderOpts), ctypes.sizeof(_lib.RenderOut)]
    from envidr_b200 import density, epilogue, optim
    aux = (ctypes.c_uint32 * 5)()
    assert lib.envidr_abi_sizes_aux(aux) == 0
    assert list(aux) == [ctypes.sizeof(density.DensityOpts), ctypes.sizeof(optim.AdamTensor), ctypes.sizeof(epilogue.LossIn),
                         ctypes.sizeof(epilogue.LossOpts), ctypes.sizeof(_lib.SampleLog)]


def test_argument_validation_without_gpu(lib):
    # null pointers / unsupported dims are rejected before any launch, with a message
    rc = lib.envidr_near_far_from_aabb(None, None, None, 4, 0.2, None, None, None)
    assert rc == -2 and b"null" in lib.envidr_last_error()
    dummy = ctypes.c_void_p(16)
    rc = lib.envidr_hash_encode_forward(dummy, dummy, dummy, dummy, 8, 4, 2, 1, 1.0, 16, 0, None, None)
    assert rc == -1 and b"D must be 2 or 3" in lib.envidr_last_error()
    rc = lib.envidr_hash_encode_forward(dummy, dummy, dummy, dummy, 8, 3, 3, 1, 1.0, 16, 0, None, None)
    assert rc == -1 and b"C must be 1, 2, 4, or 8" in lib.envidr_last_error()
    rc = lib.envidr_ide_encode_forward(dummy, None, 0.1, 8, 6, 1.0, dummy, None)
    assert rc == -1 and b"deg_view" in lib.envidr_last_error()
    # the round-1 "next row" entry points validate the same way
    assert lib.envidr_adam_step(None, 3, 0.9, 0.99, 1e-15, 0, 0, None) == -2
    assert lib.envidr_get_rays(None, 1, None, 8, 8, None, 64, None, None, None) == -2
    assert lib.envidr_density_grid_update(None, None, None, None, 8, None, None, None, None, 0, None) == -2
    assert lib.envidr_train_loss_forward(None, None, None, None, 0, None) == -2
    assert lib.envidr_pow2_scales(None, 4, None, 4, None, 0, None, None) == -2
    rc = lib.envidr_sh_encode_forward(dummy, dummy, 8, 3, 9, None, None)
    assert rc == -1
    assert lib.envidr_near_far_from_aabb(dummy, dummy, dummy, 0, 0.2, dummy, dummy, None) == 0      # empty input is a no-op
    with pytest.raises(_lib.EnvidrError):
        _lib.check(-1, "x")
    f = _lib.Field()
    assert lib.envidr_field_pack_bytes(ctypes.byref(f)) == 0                                          # rejected description


@pytest.mark.parametrize("deg", [1, 2, 3, 4, 5])
def test_ide_tables_match_reference_module(lib, golden_dir, deg):
    """The kernel's coefficient tables (built in C++) equal the reference module's buffers (ide_encoder.py:84-96)."""
    P, lmax = 2 ** deg - 1 + deg, 2 ** (deg - 1)
    mat = np.zeros((lmax + 1, P), np.float32); sigma = np.zeros(P, np.float32); ml = np.zeros((2, P), np.int32)
    rc = lib.envidr_ide_tables(deg, mat.ctypes.data_as(ctypes.c_void_p), sigma.ctypes.data_as(ctypes.c_void_p), ml.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    from oracle import oracle as O
    ml_o, mat_o, sigma_o = O.ide_tables(deg)
    np.testing.assert_array_equal(ml, ml_o)
    np.testing.assert_array_equal(sigma, sigma_o)
    np.testing.assert_allclose(mat, mat_o, rtol=2e-7, atol=0)          # same formula in C++ double vs numpy double, then fp32
    if deg in (4, 5):
        z = np.load(os.path.join(golden_dir, "ide.npz"))
        np.testing.assert_allclose(mat, z[f"mat{deg}"], rtol=2e-7, atol=0)
    mat6 = np.zeros((40, 80), np.float32)
    assert lib.envidr_ide_tables(6, mat6.ctypes.data_as(ctypes.c_void_p), sigma.ctypes.data_as(ctypes.c_void_p), ml.ctypes.data_as(ctypes.c_void_p)) == -1


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under envidr_b200/ may reference it."""
    root = os.path.dirname(_lib.__file__)
    for dirpath, _, files in os.walk(root):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "from oracle" not in src and "import oracle" not in src and "envidr_oracle" not in src, fn
