"""world_size-2 gloo tests of the N>1 host logic on CPU: window sharding and the gradient combine rule
(per-rank gradients normalised by the GLOBAL batch, summed by all-reduce, equal the single-process gradients)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import train_torch as TT
from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, spec_from_config, stride_mask, weights
from uplift_upsample_3dhpe_b200.sharding import allreduce_gradients, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 512, 513):
        for world in (1, 2, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = UpliftUpsampleConfig.preset("h36m_81", BATCH_SIZE=4)
    spec = spec_from_config(cfg)
    w = weights.init_weights(spec, 1, perturb=True)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (4, 41, 17, 2)).astype(np.float32)
    gt = rng.normal(0, 0.3, (4, 41, 17, 3)).astype(np.float32)
    m = stride_mask.batch_stride_masks_train(41, 2, cfg.MASK_STRIDE, 4, seed=0)
    lo, hi = shard_range(4, rank, world)
    loss, g = TT.loss_and_grads(spec, w, x[lo:hi], gt[lo:hi], m[lo:hi], batch_size=4)     # GLOBAL batch size
    keys = sorted(g)
    flat = torch.from_numpy(np.concatenate([g[k].reshape(-1) for k in keys]))
    allreduce_gradients(flat, dist)
    lt = torch.tensor([loss], dtype=torch.float64)
    dist.all_reduce(lt)
    if rank == 0:
        ref_loss, ref = TT.loss_and_grads(spec, w, x, gt, m, batch_size=4)
        ref_flat = np.concatenate([ref[k].reshape(-1) for k in keys])
        out["grad_err"] = float(np.abs(flat.numpy() - ref_flat).max() / np.abs(ref_flat).max())
        out["loss_err"] = abs(float(lt) - ref_loss)
    dist.destroy_process_group()


def test_two_rank_gradient_sum_equals_single_process():
    mgr = mp.get_context("spawn").Manager()      # (fork from a multi-threaded parent can deadlock)
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["grad_err"] < 1e-10 and out["loss_err"] < 1e-12
