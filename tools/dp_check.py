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
from transformergrooveinfilling_b200 import FusedAdam, FusedSGD  # noqa: E402
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
    # ---- exchange = 'p2p': gradient sum over NVLink peer memory inside the optimizer kernel (csrc/peer_opt.cu) ----------------
    # four steps (both halves of the double-buffered exchange buffer, twice) against the single-GPU run of the concatenated batch and
    # against the NCCL form; the replicas must stay BIT-identical (every rank adds the same numbers in the same order)
    for prec, cfg, opt_name in [("bf16", G.GrooveCfg(32, 16, 512, 3, 0, 16, 27), "sgd"), ("bf16", G.GrooveCfg(256, 16, 64, 2, 0, 16, 27), "adam"),
                                ("fp32", G.GrooveCfg(32, 4, 64, 2, 2, 27, 27), "adam")]:
        n = 16 * world
        x, y = [t.cuda() for t in G.det_batch(cfg, n)]
        lo, hi = shard_bounds(n, rank, world)
        make_opt = (lambda m: FusedSGD(m, 0.07)) if opt_name == "sgd" else (lambda m: FusedAdam(m, 1e-3))
        runs = {}
        for mode in ("single", "nccl", "p2p"):
            mod, _ = build_model(cfg, device=f"cuda:{local}", dropout=0.24, precision=prec)
            mod.set_seed(5, 0, 0).train()
            opt = make_opt(mod)
            losses, g1 = [], None
            if mode == "single":
                for _ in range(4):
                    m, _ = mod.train_step(x, y, 0.38)
                    opt.step()
                    losses.append(float(m[0]))
            else:
                dp = DataParallelStep(mod, opt, 0.38, overlap=False, exchange=mode)
                assert dp.exchange == mode, (dp.exchange, mode)
                for it in range(4):
                    m = dp.step(x[lo:hi].contiguous(), y[lo:hi].contiguous(), reduce_metrics=True)
                    losses.append(float(m[0]))
                    if it == 0:
                        g1 = mod.flat_grad().detach().clone()          # the SUM over ranks in both forms
                if dp.peer is not None:
                    torch.cuda.synchronize()
                    dist.barrier()
                    dp.peer.close()
            runs[mode] = (losses, mod.flat_parameters().detach().clone(), g1)
        scale = float(runs["single"][1].abs().max())
        d_single = float((runs["p2p"][1] - runs["single"][1]).abs().max()) / scale
        d_nccl = float((runs["p2p"][1] - runs["nccl"][1]).abs().max()) / scale
        dl = max(abs(a - b) / abs(b) for a, b in zip(runs["p2p"][0], runs["single"][0]))
        # first step: same parameters, same inputs -> the two forms add the same numbers (in different orders)
        dg = float((runs["p2p"][2] - runs["nccl"][2]).abs().max() / runs["nccl"][2].abs().max())
        gathered = [torch.empty_like(runs["p2p"][1]) for _ in range(world)]
        dist.all_gather(gathered, runs["p2p"][1])
        identical = all(torch.equal(g, gathered[0]) for g in gathered)
        if rank == 0:
            print(f"DP_CHECK p2p {prec} d={cfg.d_model} dec={cfg.n_dec} {opt_name} world={world} loss_rel_diff={dl:.2e} "
                  f"param_vs_single={d_single:.2e} param_vs_nccl={d_nccl:.2e} first_step_grad_vs_nccl={dg:.2e} replicas_bit_identical={identical}")
        # after 4 steps: fp32 differs by reduction order only (Adam: sign-sensitive on rounding-noise gradients, see test_gpu_parity);
        # bf16 runs that differ by 1e-8 in a parameter part at bf16-noise level within a few steps
        tol_p = 5e-3 if prec == "bf16" else (2e-3 if opt_name == "adam" else 1e-4)
        ok = ok and dl < (2e-3 if prec == "bf16" else 2e-4) and d_single < tol_p and d_nccl < tol_p and dg < 1e-5 and identical
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 3)


if __name__ == "__main__":
    main()
