set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 30 --warmup 5 --no-extras > gpurun_out/r02c_bench_n8_64.json 2> gpurun_out/r02c_n8.err
python bench.py --steps 30 --warmup 5 --no-extras > gpurun_out/r02c_bench_n1_64_samebox.json 2> /dev/null
$TR --nproc-per-node 8 --master-port 29622 bench.py --gpus 8 --steps 15 --warmup 5 --no-extras --size 128 --batch 1024 > gpurun_out/r02c_bench_n8_128.json 2> /dev/null
for f in gpurun_out/r02c_bench_*.json; do echo $f; cut -c1-200 $f; done; tail -2 gpurun_out/r02c_n8.err
