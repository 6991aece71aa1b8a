# Multi-GPU validation + records (run under: gpurun --gpus N -- bash tools/run_8gpu_checks.sh N TAG)
N=${1:-8}; TAG=${2:-r2_c}
set -x
python -m pytest tests/test_configs_gpu.py -m gpu -x -q -k "in_library" > gpurun_out/${TAG}_pytest_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_${N}gpu.log
for cfg in bls20 bn24 bls26 kzg; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --config $cfg > gpurun_out/${TAG}_bench_${N}gpu_$cfg.json 2> gpurun_out/${TAG}_bench_${N}gpu_$cfg.err; echo "bench $cfg rc=$?"; tail -3 gpurun_out/${TAG}_bench_${N}gpu_$cfg.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench_${N}gpu_$cfg.json'))
    print('$cfg', {k:d[k] for k in ('value','ms_per_step','parity','n_gpus','scaling')}, 'e2e', round(d['e2e']['ms_per_step'],3), 'first', round(d['e2e']['first_call_ms'],3), 'pinned', round(d['e2e']['pinned_ms_per_step'],3))
except Exception as e: print('$cfg', 'no result', e)
PY
done
python bench.py --gpus $N --steps 10 --warmup 3 --config bls20 --in-library-devices > gpurun_out/${TAG}_bench_inlib${N}.json 2> gpurun_out/${TAG}_bench_inlib${N}.err; echo "inlib rc=$?"; tail -3 gpurun_out/${TAG}_bench_inlib${N}.err | cut -c1-300
python bench.py --gpus $N --steps 10 --warmup 3 --config kzg --in-library-devices > gpurun_out/${TAG}_bench_inlib${N}_kzg.json 2> gpurun_out/${TAG}_bench_inlib${N}_kzg.err; echo "inlib kzg rc=$?"; tail -3 gpurun_out/${TAG}_bench_inlib${N}_kzg.err | cut -c1-300
python bench.py --impl reference --gpus $N --steps 2 --warmup 1 --config bls20 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cat gpurun_out/${TAG}_bench_reference.json | cut -c1-400
