"""BASELINE.json configs[4]: end-to-end model_mfcc_bgru inference fed by on-device MFCC, batch 8192
sharded by clip over the GPUs of one box (1024 per GPU at 8), no collective on the data path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/cfg5_model_mfcc_bgru.py [--batch 8192] [--out gpurun_out/r2_cfg5.jsonl]

The model is the reference's Network (models/model_mfcc_bgru.py:21-37: GRU(39, 512, 2 layers, bidirectional) +
Linear(1024, 12)).  Where the reference tree is mounted ($SRFE_REFERENCE or /root/reference) the UNMODIFIED module is
imported; on the GPU box it does not exist and reference sources are never copied into this repo, so tests/twins.py's
shape twin stands in -- its seeded weights are checked here against the per-tensor checksums recorded from the real module
(tests/golden/model_mfcc_bgru_logits.npz), and its logits on the golden clips against the logits the real module produced
on the CPU.  Forward = patch_model's: PCM -> fused MFCC kernel in the GRU's [B, 51, 39] layout -> cuDNN BGRU -> fc.
Reports front-end, model and total time per step (CUDA events, max over ranks) and clips/s for the whole job, with the PCM
batch resident on the device and with the reference's handoff (a CPU float32 batch, pageable, uploaded inside the step)."""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speechrecognitionproject_b200 as S  # noqa: E402
from speechrecognitionproject_b200 import patch  # noqa: E402
from speechrecognitionproject_b200.sharding import shard_range  # noqa: E402
from tests import twins  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_cfg5.jsonl"))
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    g = np.load(os.path.join(ROOT, "tests", "golden", "model_mfcc_bgru_logits.npz"))
    mod, which = twins.load("model_mfcc_bgru")
    torch.manual_seed(int(g["seed"]))
    net = mod.Network().eval()
    for k, want in zip(g["keys"], g["checksums"]):                      # the reference module's weights, regenerated from the seed
        v = net.state_dict()[str(k)].double()
        np.testing.assert_allclose([float(v.sum()), float(v.abs().sum())], want, rtol=1e-12, atol=0, err_msg=str(k))
    net = net.cuda()
    patch.patch_model(mod)
    xg = torch.from_numpy(g["clips_i16"].astype(np.float32))            # the clips the golden logits were computed on
    with torch.no_grad():
        err = float((net(xg).cpu() - torch.from_numpy(g["logits"])).abs().max())
    assert err < 2e-4, err                                              # vs the unmodified reference module's CPU logits

    b0, b1 = shard_range(a.batch, rank, world)
    gen = torch.Generator(device=dev); gen.manual_seed(20260004 + rank)
    x = (torch.randn((b1 - b0, 16000), device=dev, generator=gen) * 3000).clamp_(-32768, 32767).round_()
    xh = x.cpu()                                                         # what the reference's DataLoader hands over (pageable)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        ms_total = timed(lambda: net(x))
        ms_fe = timed(lambda: S.mfcc(x, S.R_MFCC, layout="tf"))
        f = S.mfcc(x, S.R_MFCC, layout="tf")
        ms_model = timed(lambda: patch._bgru_tail(net, f))
        t0 = time.perf_counter()
        ms_host = timed(lambda: net(xh).sum().item())                   # H2D of the PCM + front end + model + D2H of a scalar
    if rank == 0:
        rec = {"config": "cfg5: model_mfcc_bgru inference, on-device R-MFCC, batch %d sharded x%d" % (a.batch, world),
               "n_gpus": world, "batch": a.batch, "clips_per_gpu": b1 - b0, "module": which, "steps": a.steps,
               "golden_logits_max_abs_err": err,
               "ms_front_end": ms_fe, "ms_model": ms_model, "ms_total_device_resident": ms_total,
               "clips_per_s_device_resident": a.batch / ms_total * 1e3, "front_end_share": ms_fe / ms_total,
               "ms_total_host_pageable_pcm": ms_host, "clips_per_s_host_pageable_pcm": a.batch / ms_host * 1e3,
               "gpu": torch.cuda.get_device_name(0)}
        print(json.dumps(rec))
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        with open(a.out, "a") as fh:
            fh.write(json.dumps(rec) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
