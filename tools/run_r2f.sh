mkdir -p gpurun_out/r2f; O=gpurun_out/r2f
for v in "TL_HALO_GROUPS=2" "TL_HALO_FILL=1" "TL_HALO_DOUBLE=0" "TL_HALO_SMEM_KB=226"; do echo "== $v f16x2"; env $v timeout 200 python tools/profile_layers.py cfg2_2M f16x2 2>&1 | sed -n 6,10p; done > $O/variants_f16x2.txt 2>&1; cat $O/variants_f16x2.txt
