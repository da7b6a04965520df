set -x
python -m pytest tests/test_gpu_prune.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_prune.log
python tools/profile_prune.py 10 > gpurun_out/prune_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_prune2.csv python tools/profile_prune.py 2 > gpurun_out/ncu_prune.log 2>&1
cat gpurun_out/pytest_prune.log gpurun_out/prune_plain.log
