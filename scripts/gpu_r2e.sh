cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --no-header -rf -k "persistent_pipeline" --timeout 300 2>&1 | tail -5
python scripts/pipe_timing.py 4000
for v in "MLD_PIPE_DELAY=16" "MLD_PIPE_DELAY=8" "MLD_PIPE_RING=64 MLD_PIPE_DELAY=64" "MLD_PIPE_RING=64 MLD_PIPE_DELAY=48" "MLD_PIPE_K1_GROUP=4" "MLD_PIPE_K1_GROUP=4 MLD_PIPE_DELAY=16" "MLD_PIPE_BPS=4" "MLD_PIPE_BPS=5" "MLD_PIPE_HINT=0"; do echo "== $v"; env $v python scripts/pipe_timing.py 4000 | head -8; done
