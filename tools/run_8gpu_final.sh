# Final multi-GPU records of the round (run under: gpurun --gpus N -- bash tools/run_8gpu_final.sh N TAG)
N=${1:-8}; TAG=${2:-r2_ag}
run() {  # name, extra env, bench args
  name=$1; shift; envs=$1; shift
  env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/${TAG}_bench_${N}gpu_$name.json 2> gpurun_out/${TAG}_bench_${N}gpu_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_${N}gpu_$name.json').read().strip().splitlines()[-1])
    print('$name', {k:d[k] for k in ('value','ms_per_step','parity','n_gpus','scaling')}, 'e2e', round(d['e2e']['ms_per_step'],3), 'first', round(d['e2e']['first_call_ms'],3), 'pinned', round(d['e2e']['pinned_ms_per_step'],3))
except Exception as e: print('$name', 'no result', e)
PY
}
nproc
run bls20 "X=1" --config bls20
run bls20_stage2 "ZKB200_STAGE_THREADS=2" --config bls20
run bls20_nostaging "ZKB200_NO_STAGING=1" --config bls20
run bn24 "X=1" --config bn24
run kzg "X=1" --config kzg
run bls26 "X=1" --config bls26
python -m pytest tests/test_configs_gpu.py -m gpu -x -q -k "in_library" > gpurun_out/${TAG}_pytest_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest_${N}gpu.log
python bench.py --gpus $N --steps 10 --warmup 3 --config bls20 --in-library-devices > gpurun_out/${TAG}_bench_inlib${N}.json 2> gpurun_out/${TAG}_bench_inlib${N}.err; echo "inlib rc=$?"; cut -c1-250 gpurun_out/${TAG}_bench_inlib${N}.json
