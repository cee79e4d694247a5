# 2-GPU A/B of the tile-Cholesky CTA share (4 = previous default, 6 = default for N <= 1024): fusion phases of the bench
for sh in 4 6; do
XB_CHOL_SHARE=$sh timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | head -1 > gpurun_out/share${sh}_2gpu.json
python - <<P
import json
d=json.load(open("gpurun_out/share${sh}_2gpu.json"))
print("share $sh", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ci", json.dumps(d.get("ci"))[:300], "mm", json.dumps(d.get("multi_uav_msckf"))[:500])
P
done
