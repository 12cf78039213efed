set -x
O=gpurun_out/r5
mkdir -p $O
timeout 600 python tools/e2e_probe.py > $O/e2e_probe.log 2>&1; tail -2 $O/e2e_probe.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.log 2>&1; tail -2 $O/bench_n2.log
