set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/r02b_bench_n8_64.json 2> gpurun_out/r02b_n8_64.err
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r02b_bench_n1_64_samebox.json 2> /dev/null
$TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --width 2 > gpurun_out/r02b_bench_n8_64_width2.json 2> gpurun_out/r02b_n8_w2.err
python bench.py --steps 10 --warmup 3 --no-extras --width 2 > gpurun_out/r02b_bench_n1_64_width2_samebox.json 2> /dev/null
$TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --steps 8 --warmup 3 --no-extras --width 2 --size 128 --batch 512 > gpurun_out/r02b_bench_n8_128_width2.json 2> gpurun_out/r02b_n8_128w2.err
python bench.py --steps 8 --warmup 3 --no-extras --width 2 --size 128 --batch 512 > gpurun_out/r02b_bench_n1_128_width2.json 2> /dev/null
for f in gpurun_out/r02b_bench_*.json; do echo $f; cut -c1-200 $f; done; tail -3 gpurun_out/r02b_n8_64.err
