#!/bin/bash
# bash profiles/scale_run.sh N "<workload args>" tag   (under gpurun --gpus N): one strong-scaling point
N=$1; ARGS=$2; TAG=$3
OUT=gpurun_out
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-replicas $ARGS 2> $OUT/scale_${TAG}_n$N.err | grep '^{' > $OUT/scale_${TAG}_n$N.json
tail -c 300 $OUT/scale_${TAG}_n$N.err
python - <<P
import json
d=json.load(open("$OUT/scale_${TAG}_n$N.json"))
ins=d["roofline"]["in_situ"]
print("$TAG N=$N value %.1f e2e %.1f ms/step %.4f true_res %.2e cycles %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["true_residual"], d["e2e"]["step"]))
acc={}
for k,v in ins.get("exclusive_us",{}).items():
    if "." in k.split()[-1]: continue
    key=" ".join(k.split()[:1])+" "+("exchange" if "halo_exchange" in k else "compute")
    acc[key]=acc.get(key,0)+v
print("  iteration_us %.1f" % ins.get("iteration_us",0), {k: round(v,1) for k,v in acc.items()})
P
