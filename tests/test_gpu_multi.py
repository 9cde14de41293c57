"""Multi-GPU (needs >= 2 GPUs on the box; skipped otherwise): the NVLink peer mailbox that replaces the NCCL
all-reduce of the batch-summed spectrum (csrc/peer.cuh), and the shared period search built on it."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT / "flow-timesnet_b200"))
    import torch.distributed as dist
    import flowtimes_synth as syn
    from timesnet_forecast import _native as nv
    from timesnet_forecast.models.timesnet import FFTPeriodSelector, TimesBlock
    from timesnet_forecast.parallel import local_period_search, shard_batch, share_period_search
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        res = {}
        wl = syn.WORKLOADS["elec"]
        B = 8
        x_full = syn.white_features(B, wl.T, wl.d_model, seed=5).to(torch.bfloat16)
        lo, hi = shard_batch(B, rank, world)
        x = x_full[lo:hi].to(dev)
        torch.manual_seed(0)
        blk = TimesBlock(wl.d_model, [list(k) for k in wl.kernel_set], 0.0, "gelu", d_ff=wl.ff,
                         bottleneck_ratio=wl.bottleneck_ratio).to(dev).eval()
        sel = FFTPeriodSelector(wl.k_periods, wl.T, 1)
        object.__setattr__(blk, "period_selector", sel)
        # 1. the raw exchange: rank-ordered sum, several calls in a row (slot / epoch protocol), bit-identical everywhere
        share_period_search(sel, transport="peer")
        assert sel.peer_comm is not None
        sums = []
        for it in range(5):
            g = torch.Generator().manual_seed(100 * it)
            rows = torch.randn(world, 170, generator=g)
            v = rows[rank].clone().to(dev)
            sel.peer_comm.all_reduce(v)
            want = torch.zeros(170)
            for q in range(world):
                want = want + rows[q]
            assert torch.equal(v.cpu(), want), f"call {it}: peer all-reduce differs from the rank-ordered sum"
            sums.append(v.cpu())
        # 2. shared search over the mailbox == search of the full batch on one rank
        out_peer = blk(x).clone()
        res["periods_peer"] = sel.last_selected_periods.tolist()
        res["freq_peer"] = sel.last_frequency_indices.tolist()
        plan_peer = blk._last_plan.plan_dev.cpu()[:804].clone()
        # 3. same through the NCCL transport (two library calls): identical plan up to summation order, identical periods
        share_period_search(sel, transport="nccl")
        assert sel.peer_comm is None
        out_nccl = blk(x).clone()
        res["periods_nccl"] = sel.last_selected_periods.tolist()
        res["block_equal"] = bool(torch.equal(out_peer, out_nccl))
        res["block_maxdiff"] = float((out_peer.float() - out_nccl.float()).abs().max())
        # 4. rank-local search (the default) on the FULL batch: the reference's single-process answer
        local_period_search(sel)
        blk(x_full.to(dev))
        res["periods_full"] = sel.last_selected_periods.tolist()
        res["plan_peer"] = plan_peer
        torch.cuda.synchronize()
        torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_peer_mailbox_shared_search_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    for o in outs:
        assert o["periods_peer"] == o["periods_full"], "sharded search over the mailbox != full-batch search"
        assert o["periods_nccl"] == o["periods_full"]
        assert o["block_maxdiff"] < 2e-2
    assert torch.equal(outs[0]["plan_peer"], outs[1]["plan_peer"]), "ranks hold different plans"
