# runs the 64k-element mm10 check with the default library and every build/variants/*.so
python -m pytest tests/test_gpu_mm10.py -q -x 2>&1 | tail -2
for lib in "" build/variants/*.so; do
  echo "== ${lib:-default}"
  W3D_ENGINE_LIB=$lib python bench.py --workload hex8_64k_mm10 --steps 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['stress_pass'], d['plastic_check'])"
done
