cd "${GRAFT_REPO_ROOT:-.}"
./build/pcie_probe
nproc; lscpu | grep -i "model name\|^CPU(s)\|numa" | head -5
timeout 900 python -m pytest tests/test_shim_cpp.py tests/test_semantic_plane.py -m gpu -q --no-header -rf --timeout 600 2>&1 | tail -6
