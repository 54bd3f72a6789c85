"""Data-parallel training on the CUDA path: N ranks (one process per GPU, library-owned NCCL communicator, bucketed
all-reduce overlapped with the backward pass through uu_train_step) must equal one rank on the whole batch — loss,
gradients after the first step and weights after three AdamW steps.  Needs >= 2 GPUs (skipped otherwise); run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`."""
import os
import socket
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

GLOBAL_B = 16      # 8 windows x 41 tokens = 328 rows per rank: the tensor-core GEMMs (>= 256 rows) engage on both sides
STEPS = 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(cfg, spec):
    from uplift_upsample_3dhpe_b200 import stride_mask
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, (GLOBAL_B, spec.n_tok, 17, 2)).astype(np.float32)
    gt = rng.normal(0, 0.3, (GLOBAL_B, spec.n_tok, 17, 3)).astype(np.float32)
    m = stride_mask.batch_stride_masks_train(spec.n_tok, cfg.SEQUENCE_STRIDE, cfg.MASK_STRIDE, GLOBAL_B, seed=0)
    return x, gt, m


def _run(rank, world, port, out_dir, math):
    import torch.distributed as dist
    from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, spec_from_config, weights
    from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer
    from uplift_upsample_3dhpe_b200.sharding import shard_range
    from uplift_upsample_3dhpe_b200.train import Trainer
    torch.cuda.set_device(rank)
    if world > 1:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = UpliftUpsampleConfig.preset("h36m_81", BATCH_SIZE=GLOBAL_B)
    spec = spec_from_config(cfg)
    w0 = weights.init_weights(spec, 1, perturb=True)
    x, gt, m = _data(cfg, spec)
    lo, hi = shard_range(GLOBAL_B, rank, world)
    model = build_uplift_upsample_transformer(cfg, device=rank, precision="fp32", weights=w0)
    tr = Trainer(model, cfg, droppath=False, math=math)
    if world > 1:
        tr.init_comm(dist)
        assert model._lib.uu_comm_world_size(model._h) == world
    xd, gd, md = (torch.from_numpy(a[lo:hi]).cuda() for a in (x, gt, m))
    losses, g1 = [], None
    for it in range(STEPS):
        loss = tr.train_step(xd, gd, md)
        torch.cuda.synchronize()
        losses.append(float(loss.item()))
        if it == 0:
            g1 = tr.get_grads()
    if rank == 0:
        got = model.get_weights()
        np.savez(os.path.join(out_dir, f"w{world}.npz"), losses=np.array(losses),
                 **{f"g|{k[0]}|{k[1]}": v for k, v in g1.items()}, **{f"w|{k[0]}|{k[1]}": v for k, v in got.items()})
    model.close()
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("math", ["fp32", "tf32"])
def test_n_rank_step_equals_one_rank_step(math):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_run, args=(1, 0, d, math), nprocs=1, join=True)
        mp.spawn(_run, args=(world, _free_port(), d, math), nprocs=world, join=True)
        a, b = np.load(os.path.join(d, "w1.npz")), np.load(os.path.join(d, f"w{world}.npz"))
        # the sum over ranks re-associates fp32 additions of the per-window gradients: 1e-4 of the tensor's scale; the
        # tensor-core mode additionally re-tiles its split-K reductions with the local row count
        gtol = 1e-4 if math == "fp32" else 2e-3
        assert np.allclose(a["losses"], b["losses"], rtol=1e-5 if math == "fp32" else 1e-4, atol=0)
        # (key biases have a mathematically zero gradient: compare against a floor relative to the largest gradient)
        floor = 1e-6 * max(np.abs(a[k]).max() for k in a.files if k.startswith("g|"))
        for k in a.files:
            if k.startswith("g|"):
                assert np.abs(a[k] - b[k]).max() <= gtol * np.abs(a[k]).max() + floor, k
        w0 = {k: v for k, v in a.items() if k.startswith("w|")}
        for k in w0:
            # three steps move a weight by ~3 lr = 1.2e-4; weights must agree far inside that (key biases: gradient is
            # round-off noise that Adam normalises to +-lr, excluded as in test_three_adamw_steps_match_oracle)
            if k.endswith("|5") and "block" in k:
                continue
            diff = np.abs(a[k] - b[k])
            if math == "fp32":
                assert diff.max() < 2e-5, k
            else:
                # tf32 mode: the split-K tiling of the tensor-core weight gradients follows the local row count, so the
                # TF32-rounded partial sums differ between 1 and 2 ranks; where a gradient entry is dominated by that
                # rounding Adam's m / sqrt(v) turns the difference into +-lr per step (3 steps x 4e-5 x 2 at most).
                # Bulk of the entries must still agree far inside one step: 99 % within half a step (measured: the 99.9 %
                # quantile of the worst tensor sits at 2.15e-5 = half a step, the maximum at 1.1e-4).
                assert diff.max() <= 2.5e-4 and np.quantile(diff, 0.99) < 2e-5, (k, diff.max(), np.quantile(diff, 0.99))
