# usage: bash tools/_final_scale.sh N   -- the driver's multi-GPU launch of bench.py, N ranks on one node
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r01_final_scale_n$N.json 2> gpurun_out/r01_final_scale_n$N.err
tail -c 400 gpurun_out/r01_final_scale_n$N.json
