mkdir -p gpurun_out/r2c; O=gpurun_out/r2c
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12000 --csv --log-file $O/train_launches.csv python tools/profile_train.py 2 tf32 2 > $O/ncu_train.log 2>&1; tail -2 $O/ncu_train.log | cut -c1-200
python tools/summarise_launches.py $O/train_launches.csv > $O/train_launches_summary.txt; head -32 $O/train_launches_summary.txt | cut -c1-190
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $O/step_launches.csv python tools/profile_step.py cfg2_2M f16x2 > $O/ncu_step.log 2>&1; tail -2 $O/ncu_step.log | cut -c1-200
python tools/summarise_launches.py $O/step_launches.csv > $O/step_launches_summary.txt; head -45 $O/step_launches_summary.txt | cut -c1-190
