set -x
for n in 2 4 8; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $n --steps 100 --warmup 5 2>gpurun_out/r2_scale.err | grep '^{' > gpurun_out/r2_scale_${n}gpu.json
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $n --steps 100 --warmup 5 --bn-sync 2>>gpurun_out/r2_scale.err | grep '^{' > gpurun_out/r2_scale_${n}gpu_bnsync.json
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29516 tests/probe_imagenet_scaling.py 30 bf16 2>>gpurun_out/r2_scale.err | grep '^{' > gpurun_out/r2_imagenet_${n}gpu.json
done
python bench.py --gpus 1 --steps 100 --warmup 5 2>>gpurun_out/r2_scale.err | grep '^{' > gpurun_out/r2_scale_1gpu.json
python tests/probe_imagenet_scaling.py 30 2>>gpurun_out/r2_scale.err | grep '^{' > gpurun_out/r2_imagenet_1gpu.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_scale_*gpu*.json"))+sorted(glob.glob("gpurun_out/r2_imagenet_*gpu.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("ms_per_step", d.get("ms_per_pair")), d.get("value", d.get("pairs_per_s")))
    except Exception as e:
        print(f, "ERR", e)
PY
