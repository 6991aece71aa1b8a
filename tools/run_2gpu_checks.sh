set -x
python -m pytest tests/test_configs_gpu.py -m gpu -x -q -k "in_library" > gpurun_out/r2_b_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_b_pytest_2gpu.log
for cfg in bls20 bn24 kzg; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --config $cfg > gpurun_out/r2_b_bench_2gpu_$cfg.json 2> gpurun_out/r2_b_bench_2gpu_$cfg.err; echo "bench $cfg rc=$?"; tail -3 gpurun_out/r2_b_bench_2gpu_$cfg.err; python -c "
import json,sys
d=json.load(open('gpurun_out/r2_b_bench_2gpu_$cfg.json'))
print({k:d[k] for k in ('value','ms_per_step','parity','n_gpus','scaling')}, d['e2e']['ms_per_step'], d['e2e']['first_call_ms'], d['e2e']['pinned_ms_per_step'])"
done
python bench.py --gpus 2 --steps 5 --warmup 3 --config bls20 --in-library-devices > gpurun_out/r2_b_bench_inlib2.json 2> gpurun_out/r2_b_bench_inlib2.err; echo "inlib rc=$?"; tail -3 gpurun_out/r2_b_bench_inlib2.err; cat gpurun_out/r2_b_bench_inlib2.json | head -c 1500
