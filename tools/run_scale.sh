#!/bin/bash
# Round-2 multi-GPU records of the two BASELINE.json configs that name N > 1 (cfg3: Goddard 30+30 x 4096 per GPU,
# sharded 1 -> 8; cfg5: low-thrust-128 x 1024 per GPU = 8192 on 8 GPUs with the NCCL gather):
#   gpurun --gpus N -- bash tools/run_scale.sh N      -> gpurun_out/r2_scale_<workload>_n<N>.json
N=${1:-2}
mkdir -p gpurun_out
for WL in goddard_knot30x2 lowthrust128; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --workload $WL --steps 10 --warmup 3 --no-extras --no-e2e-variants --e2e-steps 3 --no-cpu-baseline \
      > gpurun_out/r2_scale_${WL}_n${N}.json 2> gpurun_out/r2_scale_${WL}_n${N}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --workload $WL --steps 10 --warmup 3 --no-extras --no-e2e-variants --e2e-steps 3 \
      > gpurun_out/r2_scale_${WL}_n${N}.json 2> gpurun_out/r2_scale_${WL}_n${N}.err
  fi
  echo "== $WL N=$N rc=$?"; tail -c 1500 gpurun_out/r2_scale_${WL}_n${N}.json; tail -3 gpurun_out/r2_scale_${WL}_n${N}.err
done
