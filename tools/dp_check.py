"""Run under torchrun on >= 2 GPUs: the data-parallel step (shards + NCCL all-reduce + 1/world in the
optimizer) equals the single-GPU step on the concatenated batch — same loss, same dropout masks (keyed by
global sequence index), same updated parameters up to fp32 reduction order."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import groove_oracle as G  # noqa: E402
from _util import build_model  # noqa: E402
from transformergrooveinfilling_b200 import FusedSGD  # noqa: E402
from transformergrooveinfilling_b200.dp import DataParallelStep, shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [("fp32", 2e-5, G.GrooveCfg(32, 16, 512, 3, 0, 16, 27), True), ("bf16", 2e-5, G.GrooveCfg(32, 16, 512, 3, 0, 16, 27), True),
             ("bf16", 2e-5, G.GrooveCfg(32, 16, 512, 3, 0, 16, 27), False), ("bf16", 2e-5, G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), True),
             ("fp32", 2e-5, G.GrooveCfg(32, 4, 64, 2, 2, 27, 27), True)]
    for prec, tol, cfg, overlap in cases:
        n = 16 * world
        x, y = [t.cuda() for t in G.det_batch(cfg, n)]
        ref, _ = build_model(cfg, device=f"cuda:{local}", dropout=0.24, precision=prec)
        ref.set_seed(5, 0, 0).train()
        ropt = FusedSGD(ref, 0.07)
        m_ref, _ = ref.train_step(x, y, 0.38)
        ropt.step()
        mod, _ = build_model(cfg, device=f"cuda:{local}", dropout=0.24, precision=prec)
        mod.set_seed(5, 0, 0).train()
        dp = DataParallelStep(mod, FusedSGD(mod, 0.07), 0.38, overlap=overlap, bucket_bytes=64 * 1024)
        assert (dp.groups is not None and len(dp.groups) >= 2) == overlap
        lo, hi = shard_bounds(n, rank, world)
        m = dp.step(x[lo:hi].contiguous(), y[lo:hi].contiguous(), reduce_metrics=True)
        dl = abs(float(m[0]) - float(m_ref[0])) / abs(float(m_ref[0]))
        dpar = float((mod.flat_parameters() - ref.flat_parameters()).abs().max() / ref.flat_parameters().abs().max())
        if rank == 0:
            print(f"DP_CHECK {prec} d={cfg.d_model} dec={cfg.n_dec} overlap={overlap} groups={len(dp.groups or [])} world={world} "
                  f"loss_rel_diff={dl:.2e} param_rel_diff={dpar:.2e}")
        ok = ok and dl < tol * 10 and dpar < 1e-4
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 3)


if __name__ == "__main__":
    main()
