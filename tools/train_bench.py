"""BASELINE config 5: the Stage-1 L1 training step (LR generation, forward, L1, backward, gradient all-reduce, clip +
AdamW), batch 32 per GPU, random-init T1 weights, data parallel over NCCL when launched under torchrun.
    python tools/train_bench.py [batch_per_gpu] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from fsr_b200 import sharding
from oracle import weights

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.distributed.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = dict(num_groups=6, blocks_per_group=10)
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(weights.make_state_dict(0, "T1", **cfg)); m = m.to(dev).train()
st = fsr_b200.Stage1Step(m)
g = torch.Generator(device=dev).manual_seed(10 + rank)
hrs = [torch.rand(B, 3, 256, 256, device=dev, generator=g) for _ in range(4)]
ev = lambda: torch.cuda.Event(enable_timing=True)
losses = []
for i in range(2):
    losses.append(st.step(hrs[i % 4])[0])
torch.cuda.synchronize(); sharding.barrier()
e0, e1 = ev(), ev()
e0.record()
for i in range(steps):
    loss, norm = st.step(hrs[i % 4]); losses.append(loss)
e1.record(); torch.cuda.synchronize(); sharding.barrier()
ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev) / steps
# breakdown of one step on rank 0 (events on the launch stream)
from fsr_b200 import data, training
marks = [ev() for _ in range(6)]
hr = hrs[0]
import time
marks[0].record(); h0 = time.perf_counter(); lr_img, _ = data.lr_from_hr_float(hr)
sr, ws = m._forward_train(lr_img); h1 = time.perf_counter(); marks[1].record()
loss2, dsr = training.l1_loss(sr, hr); marks[2].record()
h2 = time.perf_counter(); grads = m._backward(lr_img, dsr, ws); h3 = time.perf_counter(); marks[3].record()
training.allreduce_mean_(grads); marks[4].record()
st.opt.step(grads); m.mark_parameters_updated(); marks[5].record()
torch.cuda.synchronize()
if rank == 0:
    names = ["LR + forward (train, activations kept)", "L1 loss + gradient", "backward", "all-reduce", "norm + clip + AdamW"]
    print(f"C5 Stage-1 step, batch {B}/GPU x {world} GPU(s): {ms:.1f} ms/step = {B * world / ms * 1e3:.0f} images/s; "
          f"loss {losses[0].item():.5f} -> {losses[-1].item():.5f}, grad norm {norm.item():.4f}")
    for n, a, b in zip(names, marks[:-1], marks[1:]):
        print(f"   {n:42s} {a.elapsed_time(b):8.2f} ms")
    print(f"   host time to ISSUE the forward {1e3 * (h1 - h0):.2f} ms, the backward {1e3 * (h3 - h2):.2f} ms "
          f"(the GPU queue is empty when the forward starts: its device time above includes that)")
if world > 1:
    torch.distributed.destroy_process_group()
