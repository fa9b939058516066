# round-end single-GPU measurement set: both arms exactly as the driver launches them, the mm10 and
# mm05 workloads, the ncu launch list of the same default command, and one full capture of each
# mm01 kernel (256k-element mesh)
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r01_final_reference_arm.json 2> gpurun_out/r01_final_reference_arm.err
python bench.py --gpus 1 > gpurun_out/r01_final_bench_n1.json 2> gpurun_out/r01_final_bench_n1.err
python bench.py --workload hex8_1m_mm10 --steps 5 > gpurun_out/r01_final_bench_mm10_1m.json 2> gpurun_out/r01_final_bench_mm10_1m.err
for w in hex8_256k_mm05_fa hex8_256k_mm05_gp; do
  python bench.py --workload $w --steps 8 > gpurun_out/r01_final_bench_$w.json 2> gpurun_out/r01_final_bench_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r01_final_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eps_sig_ifv -s 6 -c 1 -o gpurun_out/fin2_stress -f python bench.py --workload hex8_256k_mm01 --steps 2 --warmup 1 --no-cpu > gpurun_out/fin2_stress.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tanstf -s 3 -c 1 -o gpurun_out/fin2_tanstf -f python bench.py --workload hex8_256k_mm01 --steps 2 --warmup 1 --no-cpu > gpurun_out/fin2_tanstf.log 2>&1
tail -c 600 gpurun_out/r01_final_bench_n1.json; echo; tail -c 300 gpurun_out/r01_final_reference_arm.json
