#!/bin/bash
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-sc16 --no-cpu --no-sustained --no-others > gpurun_out/r02_n8x_$name.json 2> gpurun_out/r02_n8x_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_n8x_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"]/1e6,1), "M frames/s", round(d["ms_per_step"],4), "ms/step; kernel", round(d["roofline"]["launch_ms"],4))
except Exception as e:
    print("$name failed", e)
PY
}
run default A=1
run p2p1 NCCL_MAX_P2P_NCHANNELS=1
run p2p2_res2 NCCL_MAX_P2P_NCHANNELS=2 DOA_SMS_RESERVE=2
run p2p4_res4 NCCL_MAX_P2P_NCHANNELS=4 DOA_SMS_RESERVE=4
run serial DOA_PIPELINE=0
