"""
Multi-GPU tests (need >= 2 GPUs; skipped otherwise): the fused all-gather
(rvl_loglike_dev_scatter + symmetric memory) against the NCCL all-gather, world size 2.
Run with: gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    from evidence_b200 import synth
    from evidence_b200.multigpu import FusedGatherLikelihood, ShardedLikelihood
    from evidence_b200.rvmodel import RVModel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        case = synth.make_case(2, n_epochs=300)
        model = RVModel(case.fixedpardict, case.datadict(), case.parnames, device=rank)
        rows = 1000
        theta = torch.from_numpy(case.draw_theta(rows, seed=50 + rank)).cuda()
        nccl = ShardedLikelihood(lambda blk: model.log_likelihood_device(blk), case.ndim)
        want = nccl.evaluate_local(theta).cpu().numpy()
        outs = []
        for signal in ("flags", "barrier"):
            fused = FusedGatherLikelihood(model, rows, signal=signal)
            for _ in range(5):  # alternating buffers
                outs.append(fused.evaluate_local(theta).clone())
            # a smaller block through the same buffers
            part = fused.evaluate_local(theta[:333].contiguous()).clone()
            # (another batch size picks another summation tree: last-bits differences)
            assert torch.allclose(part[rank * 333:(rank + 1) * 333],
                                  torch.from_numpy(want[rank * rows:rank * rows + 333]).cuda(),
                                  rtol=0.0, atol=5e-10)
        torch.cuda.synchronize()
        ret[rank] = (want, [o.cpu().numpy() for o in outs])
        model.close()
    finally:
        dist.destroy_process_group()


def test_fused_gather_equals_nccl_all_gather():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    want0, outs0 = ret[0]
    want1, outs1 = ret[1]
    assert np.array_equal(want0, want1) and want0.shape == (2000,)
    for o in outs0 + outs1:
        assert np.array_equal(o, want0)
