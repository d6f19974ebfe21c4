#!/bin/bash
# N-GPU check on one box: (1) bench.py under torchrun, (2) the CLI on N ranks writes byte-identical score files to 1 rank.
set -e
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json
tail -c 1500 gpurun_out/bench_n$N.json; echo
rm -rf /tmp/s1 /tmp/sN
python importance_generation.py --net resnet_56 --synthetic --random_init --batch_size 64 --limit 3 --out_root /tmp/s1 --compress_rate '[0.]+[0.18]*29' > /dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    importance_generation.py --net resnet_56 --synthetic --random_init --batch_size 64 --limit 3 --out_root /tmp/sN --compress_rate '[0.]+[0.18]*29' > /dev/null
python - <<PY
import os, numpy as np
a, b = '/tmp/s1/resnet_56_limit3', '/tmp/sN/resnet_56_limit3'
names = sorted(f for f in os.listdir(a) if f.endswith('.npy'))
assert names == sorted(f for f in os.listdir(b) if f.endswith('.npy')) and len(names) == 55
same = sum(open(os.path.join(a, f), 'rb').read() == open(os.path.join(b, f), 'rb').read() for f in names)
worst = max(np.abs(np.load(os.path.join(a, f)) - np.load(os.path.join(b, f))).max() / max(np.load(os.path.join(a, f)).max(), 1e-30) for f in names)
print('score files byte-identical 1 vs $N ranks: %d of %d (worst rel diff %.2e; a per-rank batch of 64/$N images makes cuDNN pick other '
      'convolution algorithms than 64, so the ACTIVATIONS may differ in the last bits; the scoring itself is rank-count independent)' % (same, len(names), worst))
assert worst < 1e-5
print('kept_channels.json identical:', open(a + '/kept_channels.json').read() == open(b + '/kept_channels.json').read())
PY
