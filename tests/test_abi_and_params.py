"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, parses the reference's yaml like DepthEstimatorParameters::fromFile, and fails loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes as C
import re
from pathlib import Path

import pytest

import oracle_lib as O
from mono_lidar_depth_b200 import DepthEstimator, DepthEstimatorParameters, MldError, _capi

ROOT = Path(__file__).resolve().parent.parent


def _declared_functions(header="mld_c_api.h"):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mld_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libmld_cuda.so does not export {n}"
        assert n in _capi.SYMBOLS, f"{n} declared in mld_c_api.h but not bound in _capi.py"
    assert set(_capi.SYMBOLS) == set(names)


def test_synth_library_exports_every_declared_symbol_and_needs_no_cuda():
    """include/mld_synth.h -> libmld_synth.so: the host generators (CPU arm of bench.py) never map the product library."""
    lib = _capi.load_synth()
    names = _declared_functions("mld_synth.h")
    assert len(names) >= 5
    for n in names:
        assert hasattr(lib, n), f"libmld_synth.so does not export {n}"
    assert set(_capi.SYNTH_SYMBOLS) == set(names)
    import subprocess

    deps = subprocess.run(["ldd", str(_capi.SYNTH_LIB_PATH)], capture_output=True, text=True).stdout
    assert "cuda" not in deps.lower() and "mld_cuda" not in deps


def test_params_layout_matches_oracle_and_header():
    lib = _capi.load()
    assert lib.mld_sizeof_params() == C.sizeof(_capi.MldParams)
    a, b = _capi.MldParams(), O.default_params()
    lib.mld_default_params(C.byref(a))
    assert bytes(a) == bytes(b)  # product defaults == oracle defaults == DepthEstimatorParameters.h member initialisers
    assert a.viewray_plane_orthoganality_treshold == 1.0  # `{01}` octal literal upstream


def test_yaml_loader_matches_reference_file(tmp_path):
    yaml = tmp_path / "parameters.yaml"
    # the values of /root/reference/monolidar_fusion/parameters.yaml (flat OpenCV yaml), abridged to the hot-path keys
    yaml.write_text(
        "%YAML:1.0\n\n"
        "neighbor_search_mode: 0 # 0-> pixels\n"
        "pixelarea_search_witdh: 6\npixelarea_search_height: 9\nradiusSearch_count_min: 1\n"
        "do_use_histogram_segmentation: 1\nhistogram_segmentation_bin_witdh: 0.3 # in meters\n"
        "histogram_segmentation_min_pointcount: 3\ndo_use_depth_segmentation: 1 # takes some time\n"
        "treshold_depth_enabled: 1\ntreshold_depth_mode: 0 \ntreshold_depth_max: 100\ntreshold_depth_min: 0\n"
        "treshold_depth_local_enabled: 1\ntreshold_depth_local_mode: 0 \ntreshold_depth_local_valuetype: 1 \n"
        "treshold_depth_local_value: 0.5\ndo_use_PCA: 0\npca_debug: 0.01\npca_treshold_3_abs_min: 0.005\n"
        "pca_treshold_3_2_rel_max: 15\npca_treshold_2_1_rel_min: 1.5\ndo_use_ransac_plane: 1\n"
        "ransac_plane_distance_treshold: 0.3\nransac_plane_max_iterations: 10000\nransac_plane_probability: 0.999\n"
        "ransac_plane_use_refinement: 1\nransac_plane_refinement_treshold: 10.2\nransac_plane_point_distance_treshold: 0.2\n"
        "ransac_plane_use_camx_treshold: 0\nransac_plane_treshold_camx: 2.0\nplane_estimator_use_triangle_maximation: 0 \n"
        "plane_estimator_use_leastsquares: 0\nplane_estimator_use_mestimator: 1\nplane_estimator_z_x_min_relation: 0\n"
        "do_use_cut_behind_camera: 1\ndo_use_triangle_size_maximation: 1\ndo_check_triangleplanar_condition: 1\n"
        "triangleplanar_crossnorm_treshold: 0.1\nviewray_plane_orthoganality_treshold: 0.03\n"
    )
    p = DepthEstimatorParameters()
    p.fromFile(str(yaml))
    ref = DepthEstimatorParameters.reference_yaml()
    d, r = p.as_dict(), ref.as_dict()
    # keys the yaml does not hold read as 0 through cv::FileStorage (fromFile has no defaults)
    assert d["ransac_plane_min_z"] == 0.0 and d["ransac_plane_max_z"] == 0.0 and d["set_all_depths_to_zero"] == 0
    assert d["do_use_depth_segmentation"] == 1  # the shipped value; reference_yaml() forces it to 0
    assert d["pca_debug"] == 0  # (int) of a real node rounds 0.01 to 0
    for k in d:
        if k in ("ransac_plane_min_z", "ransac_plane_max_z", "do_use_depth_segmentation"):
            continue
        assert d[k] == r[k], k
    with pytest.raises(MldError) as ei:
        p.fromFile(str(tmp_path / "missing.yaml"))
    assert "Cant find settings file" in str(ei.value)


def test_status_names():
    lib = _capi.load()
    assert lib.mld_status_name(1) == b"Success"
    assert lib.mld_status_name(2) == b"RadiusSearchInsufficientPoints"
    assert lib.mld_status_name(16) == b"SuccessRoad"


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must refuse to run (it never routes through the oracle)."""
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    est = DepthEstimator()
    with pytest.raises(MldError) as ei:
        est.InitConfig(DepthEstimatorParameters.reference_yaml())
    assert ei.value.code == _capi.MLD_ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_product_package_never_imports_oracle():
    """The product path may mention the oracle in comments but must never include, link, load or call it."""
    forbidden = ("mld_oracle", "orc_", "oracle_lib", "libmld_oracle", "oracle/")
    dirs = [ROOT / "mono_lidar_depth_b200", ROOT / "shim", ROOT / "include"]
    for d in dirs:
        for path in d.rglob("*"):
            if path.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp") or path.name == "Makefile":
                txt = path.read_text()
                for tok in forbidden:
                    assert tok not in txt, f"{path} references {tok}"
