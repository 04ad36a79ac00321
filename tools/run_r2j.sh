mkdir -p gpurun_out/r2j; O=gpurun_out/r2j
timeout 900 python -m pytest tests/test_gpu_trained_scale.py -m gpu -x -q -s -k "mixed or f16" > $O/pytest.log 2>&1; grep -E "mixed|f16|passed|failed|Error" $O/pytest.log | head -20
for k in 1 2; do echo "== mixed split_levels=$k"; TL_SPLIT_LEVELS=$k timeout 300 python tools/profile_layers.py cfg2_2M mixed 2>&1 | sed -n 3,14p; done > $O/layers_mixed.txt 2>&1; cat $O/layers_mixed.txt
