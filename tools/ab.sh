cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "rc=$? lines=$(wc -l < gpurun_out/bench_n8.json)"; cut -c1-400 gpurun_out/bench_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 1 --warmup 1 --nx 3840 --ny 2160 --spp 128 > gpurun_out/bench_4k_n8.json 2> gpurun_out/bench_4k_n8.err
echo "rc=$? lines=$(wc -l < gpurun_out/bench_4k_n8.json)"; cut -c1-400 gpurun_out/bench_4k_n8.json
python tools/render_once.py --nx 3840 --ny 2160 --ns 1024 --steps 1 > gpurun_out/render_4k_1024_n1.json 2>&1
cat gpurun_out/render_4k_1024_n1.json
