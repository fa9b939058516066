# profile the heaviest k_mm10 launch of the 64k-element check (launch list first, then one full capture)
true
ncu --set full --clock-control none --import-source on -k regex:k_mm10 -s ${SKIP:-4} -c 1 -o gpurun_out/mm10_cur -f python bench.py --workload hex8_64k_mm10 --steps 1 --warmup 1 --no-cpu > gpurun_out/mm10_cur.log 2>&1
ls -la gpurun_out/
