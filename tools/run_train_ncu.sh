mkdir -p gpurun_out/r2p; O=gpurun_out/r2p
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 1800 --csv --log-file $O/train_launches.csv python tools/profile_train.py 2 tf32 2 > $O/ncu_train.log 2>&1; tail -2 $O/ncu_train.log | cut -c1-200
python tools/summarise_launches.py $O/train_launches.csv > $O/train_launches_summary.txt; head -30 $O/train_launches_summary.txt | cut -c1-170
