set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "rc=$?" >> gpurun_out/r2_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rank or comm" > gpurun_out/r2_t6_comm.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err
tail -n 5 gpurun_out/r2_bench_n2.err gpurun_out/r2_t6_comm.log; head -c 1500 gpurun_out/r2_bench_n2.json; head -c 600 gpurun_out/r2_bench_ref_n2.json
