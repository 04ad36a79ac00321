mkdir -p gpurun_out/r2u; O=gpurun_out/r2u
for g in 1 2; do echo "== TL_HALO_GROUPS=$g"; TL_HALO_GROUPS=$g timeout 200 python tools/profile_layers.py cfg2_2M f16 2>&1 | sed -n 6,10p; done > $O/groups.txt 2>&1; cat $O/groups.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_halo -c 2 -o $O/halo_c32 python tools/profile_layers.py cfg2_2M f16 > $O/ncu.log 2>&1; tail -3 $O/ncu.log
timeout 300 ncu --set full --clock-control none -k regex:k_halo_build -c 1 -o $O/halo_build python tools/profile_layers.py cfg2_2M f16 > $O/ncu2.log 2>&1; tail -3 $O/ncu2.log
ls -la $O
