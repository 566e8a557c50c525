"""Where the host time of one config-5 training step goes (cProfile; not a bench number)."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pcgcv2_b200
from pcgcv2_b200 import _lib, train
pcgcv2_b200.install_shims()
import MinkowskiEngine as ME
from pcgcv2_b200.model import PCCModel

torch.manual_seed(0)
model = PCCModel().cuda().train()
bucket = train.GradBucket(model.parameters())
opt = torch.optim.Adam(model.parameters(), lr=8e-4)
c, f = train.shell_batch(0, batch=32)
c, f = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()


def step():
    x = ME.SparseTensor(features=f, coordinates=c, device="cuda")
    return float(train.train_step(model, opt, bucket, x)[0])


for _ in range(3):
    step()
torch.cuda.synchronize()
l0, t0 = _lib.launch_count(), time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"{1e3 * (time.perf_counter() - t0) / 5:.1f} ms/step, {(_lib.launch_count() - l0) / 5:.0f} libpcgc launches/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
