mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02v_fwd_gpus8.json 2> gpurun_out/r02v_fwd_gpus8.err; echo "fwd x8 rc=$?"; head -c 500 gpurun_out/r02v_fwd_gpus8.json; echo
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --mode train --steps 5 --warmup 3 > gpurun_out/r02v_train_gpus8.json 2> gpurun_out/r02v_train_gpus8.err; echo "train x8 rc=$?"; head -c 500 gpurun_out/r02v_train_gpus8.json; echo
