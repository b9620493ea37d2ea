timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu.log
timeout 400 python tools/bench_stages.py > gpurun_out/stages.json 2> gpurun_out/stages.md
