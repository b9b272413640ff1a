#!/bin/bash
# third session of round 2 (about 200 s of GPU time left): config-5 workload sample, the accelerate() path, the full GPU suite
# (incl. the new pipeline / accelerate tests), and a regression run of the default bench line
O=gpurun_out/r02_s4; mkdir -p $O
timeout 50 python bench.py --workload nu --nu-sequences 24 > $O/bench_nu24.json 2> $O/bench_nu24.err; echo "nu rc=$?"
timeout 35 python bench.py --path accelerate --no-lidal --no-cpu-baseline --no-extras --steps 10 --warmup 3 > $O/bench_accelerate.json 2> $O/bench_accelerate.err; echo "acc rc=$?"
timeout 110 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 60 python bench.py --no-cpu-baseline --steps 10 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
for n in ("bench_nu24", "bench_accelerate", "bench_default"):
    try:
        d = json.load(open(f"gpurun_out/r02_s4/{n}.json"))
        print(n, round(d["value"], 1), d["unit"], "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "lidal", d.get("lidal_frames_per_sec"),
              json.dumps(d.get("phases_ms_max_over_ranks")), json.dumps(d.get("clocks")))
    except Exception as e:
        print(n, "unreadable", repr(e))
PY
tail -c 600 $O/bench_nu24.err $O/bench_accelerate.err
