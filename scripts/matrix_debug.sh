run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/m_$tag.log 2>&1; echo "== $tag: $(grep -c '"value"' gpurun_out/m_$tag.log) ok; $(grep -o 'kernel ([^)]*) failed[^"]*\|CudaError: [^"]*' gpurun_out/m_$tag.log | head -1)"; }
run base A=1
run base2 A=1
run base3 A=1
