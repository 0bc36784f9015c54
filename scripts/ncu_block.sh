export CETPICK_BLOCK=1
CMD="python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_block' -s 2 -c 1 -o gpurun_out/r4d_block $CMD > gpurun_out/r4d_ncu.log 2>&1
ncu -i gpurun_out/r4d_block.ncu-rep --page details 2>/dev/null | grep -v "^\s*$" > gpurun_out/r4d_block_details.txt
ncu -i gpurun_out/r4d_block.ncu-rep --page source --csv 2>/dev/null > gpurun_out/r4d_block_source.csv
