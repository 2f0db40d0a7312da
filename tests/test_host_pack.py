"""Host-side record packing of the host-buffer pipeline (csrc/mld_host_pack.cpp): 16- / 32-byte records -> 12-byte xyz,
AVX-512 line squeeze where the host has it, scalar loop otherwise. CPU only: the helper is plain C++ inside libmld_cuda.so."""
import ctypes as C

import numpy as np
import pytest

from mono_lidar_depth_b200 import _capi


@pytest.mark.parametrize("stride_floats", [4, 8, 5])
def test_pack_xyz_matches_numpy(stride_floats):
    lib = _capi.load()
    lib.mld_host_pack_xyz.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_int]
    lib.mld_host_pack_xyz.restype = None
    lib.mld_host_pack_level.restype = C.c_int
    assert lib.mld_host_pack_level() in (0, 512)
    rng = np.random.default_rng(7)
    for n in (0, 1, 15, 16, 17, 63, 64, 1000, 15001):
        for dst_off, cached in ((0, 0), (1, 0), (3, 1), (5, 0), (16, 1), (0, 1)):
            src = rng.standard_normal((n + 1, stride_floats)).astype(np.float32)
            src[rng.random(n + 1) < 0.05] = np.nan
            raw = np.full(3 * n + dst_off + 64, -7.0, np.float32)
            dst = raw[dst_off:]
            lib.mld_host_pack_xyz(src.ctypes.data, 4 * stride_floats, dst.ctypes.data, n, cached)
            want = src[:n, :3].reshape(-1)
            assert np.array_equal(dst[: 3 * n].view(np.uint32), want.view(np.uint32)), (n, dst_off)
            assert np.all(raw[:dst_off] == -7.0) and np.all(dst[3 * n :] == -7.0), "wrote outside the destination"
