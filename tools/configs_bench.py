"""Measures the BASELINE.json configurations that bench.py does not print (they are parity-test cases for the
driver): C3 = one residual group (10 RCAB + group conv, the fused conv + SE + residual chain)
at batch 256; C4 = the sharded pipeline uint8 HR 256x256 -> integer LR kernel -> forward, images/s per GPU.
(C5, the Stage-1 training step, is `python bench.py --workload train` / tools/train_bench.py.)
    python tools/configs_bench.py [n_images_c4]          (one process per GPU under torchrun for N > 1)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from fsr_b200 import _lib, sharding
from oracle import weights

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.distributed.init_process_group("nccl", device_id=dev)
lib = _lib.load()
ev = lambda: torch.cuda.Event(enable_timing=True)

# ---------------------------------------------------------------- C3: one residual group, batch 256
if rank == 0:
    cfg = dict(num_groups=1, blocks_per_group=10)
    m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(weights.make_state_dict(0, "T1", **cfg)); m = m.to(dev).eval()
    x = torch.rand(256, 3, 64, 64, device=dev)
    lib.fen_profile_body(1)
    ms = []
    with torch.no_grad():
        for i in range(8):
            m(x); t = lib.fen_last_body_ms()
            if i >= 3: ms.append(t)
    lib.fen_profile_body(0)
    k_ms = sum(ms) / len(ms)
    convs = 2 * 10 + 1 + 1                      # 20 RCAB convs + group conv + conv_after_body in the same launch
    flop = convs * 2.0 * 4096 * 64 * 64 * 9 * 256
    print(f"C3 residual group (10 RCAB + group conv + conv_after_body = {convs} fused convs), batch 256: "
          f"{k_ms:.3f} ms per launch = {k_ms / convs * 1e3:.1f} us per conv layer, {flop / k_ms / 1e9:.0f} TFLOP/s "
          f"({flop / k_ms / 1e9 / 1372.1 * 100:.1f} % of the sustained bf16 peak)")
    del m, x
    torch.cuda.empty_cache()

# ---------------------------------------------------------------- C4: HR u8 -> LR kernel -> forward, sharded
total = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
b0, b1 = sharding.shard_range(total, rank, world)
cfg = dict(num_groups=6, blocks_per_group=10)
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(weights.make_state_dict(0, "T1", **cfg)); m = m.to(dev).eval()
B = 64
g = torch.Generator(device=dev).manual_seed(100 + rank)
pool = [torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(12)]  # 151 MB > L2
def run(n_img):
    done = 0; i = 0
    with torch.no_grad():
        while done < n_img:
            nb = min(B, n_img - done)
            _, lr = fsr_b200.lr_from_hr(pool[i % len(pool)][:nb], want_u8=False)
            sr = m(lr)
            done += nb; i += 1
    return sr
run(3 * B)
torch.cuda.synchronize(); sharding.barrier()
e0, e1 = ev(), ev()
e0.record(); run(b1 - b0); e1.record()
torch.cuda.synchronize(); sharding.barrier()
ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
if rank == 0:
    print(f"C4 sharded pipeline (uint8 HR 256x256 -> integer LR kernel -> forward), {total} images over {world} GPU(s), "
          f"batches of {B}: {ms:.1f} ms = {total / ms * 1e3:.0f} images/s (SR left on the device)")
if world > 1:
    torch.distributed.destroy_process_group()
