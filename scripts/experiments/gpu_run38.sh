python scripts/bench_latency.py 2>&1 | tail -1
