"""Run exactly ONE eager UNet step (BASELINE configs[1] shapes) between cudaProfilerStart/Stop, after two
warm steps -- the command the ncu captures under profiles/ were taken with:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python profiles/run_step_for_ncu.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:kv_attn -c 4 \
      -o gpurun_out/k1 python profiles/run_step_for_ncu.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200.stream_pipeline import B200StreamPipeline  # noqa: E402
from live2diff_b200.unet_step import B200UNetStep  # noqa: E402
from live2diff_b200.weights import UNetDims, random_state_dict  # noqa: E402

dev = torch.device("cuda:0")
d = UNetDims()
n_rows = int(os.environ.get("L2D_ROWS", "2"))
t_index = {1: [40], 2: [30, 40], 4: [25, 31, 37, 43]}[n_rows]
unet = B200UNetStep(random_state_dict(d, seed=0), d, n_rows, 64, 64, use_cuda_graph=False, device=dev)
pipe = B200StreamPipeline(unet, t_index)
kv = unet.prepare_cache(n_rows)
for c in kv:
    c.normal_()
pipe.prepare(torch.randn(1, 77, 768), kv)
for _ in range(48):
    pipe.schedule.advance()
x = torch.randn(1, 4, 1, 64, 64, device=dev).half()
dep = torch.randn(1, 4, 1, 64, 64, device=dev).half()
for _ in range(2):
    pipe(x, dep)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
pipe(x, dep)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step;", unet.launches_per_step, "launches")
