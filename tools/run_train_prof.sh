python tools/profile_train.py > gpurun_out/plain_train.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/profile_train.py > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/ncu_train.log
