# full ncu captures of the final kernels: mm01 stress + stiffness (256k mesh), k_mm10 (64k mesh), assembly
ncu --set full --clock-control none --import-source on -k regex:k_eps_sig_ifv -s 6 -c 1 -o gpurun_out/fin_stress -f python bench.py --workload hex8_256k_mm01 --steps 2 --warmup 1 --no-cpu > gpurun_out/fin_stress.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_assemble_nodes -s 3 -c 1 -o gpurun_out/fin_asm -f python bench.py --workload hex8_256k_mm01 --steps 2 --warmup 1 --no-cpu > gpurun_out/fin_asm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mm10 -s 6 -c 1 -o gpurun_out/fin_mm10 -f python bench.py --workload hex8_64k_mm10 --steps 1 --warmup 1 --no-cpu > gpurun_out/fin_mm10.log 2>&1
ls -la gpurun_out/*.ncu-rep
