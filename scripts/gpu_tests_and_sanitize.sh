cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 2>&1 | tail -6
bash scripts/gpu_sanitize.sh r2
