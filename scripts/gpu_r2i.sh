cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_semantic_plane.py tests/test_ransac_gpu.py -m gpu -q --no-header -rf --timeout 600 2>&1 | tail -4
export MLD_BENCH_CPU_SECONDS=0 MLD_BENCH_E2E_FRAMES=16 MLD_BENCH_NO_PARITY=1
q() { python bench.py --steps 5 --warmup 3 $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],3), {k:round(v['avg_launch_ms'],4) for k,v in d['roofline']['per_kernel'].items() if isinstance(v,dict) and v.get('avg_launch_ms')})"; }
q base
q base2
MLD_FUSE_CHUNK=384 q chunk384
MLD_FUSE_CHUNK=768 q chunk768
q road "--workload road"
q dense "--workload dense"
