SKIP_NCU=1 bash tools/gpu_round.sh r01d
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r01d/bench_n2.json 2> gpurun_out/r01d/bench_n2.err; echo "bench n2 exit $?" | tee -a gpurun_out/r01d/status.txt
tail -3 gpurun_out/r01d/bench_n2.err
head -c 600 gpurun_out/r01d/bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m pytest tests -q -m gpu -x -k distributed > gpurun_out/r01d/pytest_n2.log 2>&1; echo "pytest n2 exit $?" | tee -a gpurun_out/r01d/status.txt
