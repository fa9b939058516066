# default library and every build/variants/*.so: GPU parity tests, then the 2M mm01 bench (resident timing only)
for lib in "" build/variants/*.so; do
  echo "== ${lib:-default}"
  W3D_ENGINE_LIB=$lib python -m pytest tests -m gpu -q 2>&1 | tail -1
  W3D_ENGINE_LIB=$lib python bench.py --steps 8 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stress_pass']['ms'], d['stiffness_pass']['ms'])"
done
