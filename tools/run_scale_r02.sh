set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/r02_bench_n8_64.json 2> gpurun_out/r02_bench_n8_64.err
$TR --nproc-per-node 8 --master-port 29602 bench.py --gpus 8 --steps 15 --warmup 5 --no-extras --size 128 --batch 1024 > gpurun_out/r02_bench_n8_128.json 2> gpurun_out/r02_bench_n8_128.err
$TR --nproc-per-node 4 --master-port 29603 bench.py --gpus 4 --steps 15 --warmup 5 --no-extras --size 128 --batch 1024 > gpurun_out/r02_bench_n4_128.json 2> gpurun_out/r02_bench_n4_128.err
$TR --nproc-per-node 2 --master-port 29604 bench.py --gpus 2 --steps 15 --warmup 5 --no-extras --size 128 --batch 1024 > gpurun_out/r02_bench_n2_128.json 2> gpurun_out/r02_bench_n2_128.err
python bench.py --steps 15 --warmup 5 --no-extras --size 128 --batch 1024 > gpurun_out/r02_bench_n1_128.json 2> gpurun_out/r02_bench_n1_128.err
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r02_bench_n1_64_samebox.json 2> /dev/null
$TR --nproc-per-node 8 --master-port 29605 tests/run_dp_library_comm.py > gpurun_out/r02_dp_comm_n8.log 2>&1
for f in gpurun_out/r02_bench_n*.json; do echo $f; cut -c1-260 $f; done; tail -3 gpurun_out/r02_dp_comm_n8.log
