#!/bin/bash
N=$(nvidia-smi -L | wc -l)
run() { env "$@" timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/ddp_time.py 2>&1 | grep "^N="; }
run SC_EXCHANGE=none
run SC_EXCHANGE=allreduce
run SC_EXCHANGE=allreduce NCCL_MAX_NCHANNELS=4
run SC_EXCHANGE=allreduce NCCL_MAX_NCHANNELS=8
run SC_EXCHANGE=allreduce NCCL_ALGO=Tree
run SC_EXCHANGE=sharded
