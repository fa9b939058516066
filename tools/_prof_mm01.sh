# one full ncu capture each of the mm01 stress and stiffness kernels on the 256k-element workload
ncu --set full --clock-control none --import-source on -k regex:k_eps_sig_ifv -s ${SKIP:-6} -c 1 -o gpurun_out/mm01_stress -f python bench.py --workload hex8_256k_mm01 --steps 2 --warmup 1 --no-cpu > gpurun_out/mm01_stress.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tanstf -s ${SKIPT:-3} -c 1 -o gpurun_out/mm01_tanstf -f python bench.py --workload hex8_256k_mm01 --steps 2 --warmup 1 --no-cpu > gpurun_out/mm01_tanstf.log 2>&1
ls -la gpurun_out/
