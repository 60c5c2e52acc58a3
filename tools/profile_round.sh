set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-variants --sustain-s 0 --streams 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 416 -c 232 --csv --log-file gpurun_out/r02_ncu_launches.csv $B > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 470 -c 26 -o /tmp/layer $B > /dev/null 2>&1
python tools/ncu_summary.py /tmp/layer.ncu-rep gpurun_out/r02_ncu_layer.md
timeout 900 ncu --set full --clock-control none -k regex:"prior_tokens|roi_tc|pair_assemble|cache_fused|cache_combine|emit_kernel" -s 6 -c 6 -o /tmp/tail $B > /dev/null 2>&1
python tools/ncu_summary.py /tmp/tail.ncu-rep gpurun_out/r02_ncu_tail.md
timeout 900 ncu --set full --clock-control none -k regex:"stem_conv|maxpool|conv_gather|avgpool|gemm2" -c 60 -o /tmp/dino python tools/dino_probe.py > /dev/null 2>&1
python tools/ncu_summary.py /tmp/dino.ncu-rep gpurun_out/r02_ncu_dino.md
ls -la gpurun_out/ | head -30
