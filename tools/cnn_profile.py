"""One forward + backward of the resnet_cnn front-end inside a cudaProfilerStart/Stop range (launch list with ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsr_tf1_b200.layers import BuildContext
from avsr_tf1_b200.params import ParamStore
from avsr_tf1_b200.video import ResNetCNN

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = BuildContext()
cnn = ResNetCNN(ctx, 36, 36, 3)
ctx.store = ParamStore(ctx.specs, device='cuda', with_optimizer=True)
ctx.store.initialize(7)
N = B * 75
frames = torch.rand(N, 36, 36, 3, device='cuda') * 2 - 1
d = torch.randn(N, 128, device='cuda') * 1e-3
f = cnn.forward(frames, True); cnn.backward(d)
torch.cuda.synchronize()
torch.cuda.profiler.start()
f = cnn.forward(frames, True); cnn.backward(d)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
