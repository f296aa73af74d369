mkdir -p gpurun_out
nvidia-smi -L
timeout 120 python -m pytest tests -m gpu -q -p no:cacheprovider -k "non_current_device" > gpurun_out/r02q_gpu2_devtest.log 2>&1; echo "dev test rc=$?"; tail -2 gpurun_out/r02q_gpu2_devtest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 5 --warmup 3 > gpurun_out/r02q_train_gpus2.json 2> gpurun_out/r02q_train_gpus2.err; echo "train x2 rc=$?"; head -c 1500 gpurun_out/r02q_train_gpus2.json; echo
timeout 300 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r02q_train_gpus1.json 2> gpurun_out/r02q_train_gpus1.err; echo "train x1 rc=$?"; head -c 600 gpurun_out/r02q_train_gpus1.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_fwd_gpus2.json 2> gpurun_out/r02q_fwd_gpus2.err; echo "fwd x2 rc=$?"; head -c 400 gpurun_out/r02q_fwd_gpus2.json; echo
tail -3 gpurun_out/r02q_train_gpus2.err
