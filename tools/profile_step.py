"""One eager training step of the bench workload inside a cudaProfilerStart/Stop range, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
(see /opt/skills/guides/B200_PROFILING.md).  Not a benchmark: numbers under a profiler are never reported."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from avsr_tf1_b200 import ops  # noqa: E402
from avsr_tf1_b200.seq2seq import Seq2SeqModel  # noqa: E402
from tests.helpers import to_data_sequences  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--video-input', default='crops3888')
ap.add_argument('--attention', default=None)
ap.add_argument('--graph', default='default', choices=['default', 'parity'])
ap.add_argument('--no-tensor-cores', action='store_true')
ap.add_argument('--cuda-graph', action='store_true')
args = ap.parse_args()
ops.set_tensor_cores(not args.no_tensor_cores)
hp, batch, _ = bench.workload(args, 5, args.batch, seed=0, graph=args.graph)
batch.pop('video_u8', None)
ds = to_data_sequences(batch)
model = Seq2SeqModel(ds, 'train', hp, seed=2001)
model.use_cuda_graph = args.cuda_graph
model.feed(ds)
for _ in range(2):
    model.train_step(fetch=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model.train_step(fetch=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('launches in the profiled step:', model.launches_last_step)
