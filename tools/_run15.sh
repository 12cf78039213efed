set -x
O=gpurun_out/r15
mkdir -p $O
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.log 2>&1
  grep '^{"metric"' $O/bench_n$N.log | tail -1 | cut -c1-200
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_n1.log 2>&1
grep '^{"metric"' $O/bench_n1.log | tail -1 | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench_extra.py --workload cork --gpus 8 --steps 5 > $O/cork_n8.log 2>&1
grep '^{"metric"' $O/cork_n8.log | tail -1 | cut -c1-300
