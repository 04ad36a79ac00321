mkdir -p gpurun_out/r2n; O=gpurun_out/r2n
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_subm_probe|k_halo_build|k_hash_build|k_level_|k_emit|k_point|k_mark|k_batch|DeviceRadix|DeviceScan" -c 400 --csv --log-file $O/geom.csv python tools/profile_step.py cfg2_2M f16x2 > $O/ncu.log 2>&1
python tools/summarise_launches.py $O/geom.csv > $O/geom_summary.txt; cat $O/geom_summary.txt | cut -c1-150
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2n/geom.csv',errors='replace')) if len(r)>10]
ix={h:i for i,h in enumerate(rows[0])}
for r in rows[1:]:
    if r[ix['Metric Name']]=='gpu__time_duration.sum' and ('probe' in r[ix['Kernel Name']] or 'halo_build' in r[ix['Kernel Name']]) and int(r[ix['ID']])<60:
        print(r[ix['ID']], r[ix['Kernel Name']][:30], r[ix['Grid Size']], r[ix['Metric Value']], r[ix['Metric Unit']])
PY
