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
        # the same exchange through HOST buffers, one library call per rank (rvl_loglike_gather)
        fused = FusedGatherLikelihood(model, rows, signal="flags")
        th_pin = torch.from_numpy(case.draw_theta(rows, seed=50 + rank)).pin_memory()
        out_pin = torch.empty(world * rows, dtype=torch.float64).pin_memory()
        for _ in range(3):
            out_pin.zero_()
            fused.evaluate_local_host(th_pin.numpy(), out_pin.numpy())
            outs.append(out_pin.clone())
        pageable = np.empty(world * rows)
        fused.evaluate_local_host(np.array(th_pin.numpy()), pageable)
        outs.append(torch.from_numpy(pageable))
        # ... and through a host segment shared by the two processes (rvl_loglike_scatter_host)
        from evidence_b200.multigpu import SharedHostGather
        shared = SharedHostGather(model, rows)
        for _ in range(4):  # alternating segments
            outs.append(torch.from_numpy(shared.evaluate_local_host(th_pin.numpy()).copy()))
        shared.close()
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


def _single_device_blocks(case, theta, n_dev, device=0):
    """lnL of the contiguous row blocks a multi-device handle forms, each evaluated on ONE device."""
    from evidence_b200.multigpu import shard_bounds
    from evidence_b200.rvmodel import RVModel
    model = RVModel(case.fixedpardict, case.datadict(), case.parnames, device=device)
    out = np.empty(len(theta))
    for r in range(n_dev):
        lo, hi, _ = shard_bounds(len(theta), n_dev, r)
        if hi > lo:
            out[lo:hi] = model.log_likelihood_batch(theta[lo:hi])
    model.close()
    return out


@pytest.mark.parametrize("n_dev", [1, 2, 4])
def test_multi_device_handle_one_host_call(n_dev):
    """rvl_create_multi: one host call, rows sharded over the devices inside the library; results
    bit-identical to evaluating the same row blocks on a single device."""
    import torch
    from evidence_b200 import synth
    from evidence_b200.rvmodel import RVModel
    if torch.cuda.device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    case = synth.make_case(2, n_epochs=300)
    B = 4099  # ragged: the last block is shorter
    theta = case.draw_theta(B, seed=11)
    want = _single_device_blocks(case, theta, n_dev)
    model = RVModel(case.fixedpardict, case.datadict(), case.parnames, devices=list(range(n_dev)))
    assert model.device_count() == n_dev
    got = model.log_likelihood_batch(theta)  # pageable numpy buffers
    assert np.array_equal(got, want)
    th_pin = torch.from_numpy(theta).pin_memory()
    out_pin = torch.empty(B, dtype=torch.float64).pin_memory()
    model.log_likelihood_batch(th_pin.numpy(), out=out_pin.numpy())  # read / written in place
    assert np.array_equal(out_pin.numpy(), want)
    # fewer rows than devices, and the fused transform + likelihood through the same handle
    # (a batch of one is cut into other work items: last-bits differences in the sums)
    assert np.allclose(model.log_likelihood_batch(theta[:1]), want[:1], rtol=0.0, atol=5e-10)
    model.set_priors(case.priordict)
    U = case.draw_unit(1000, seed=3)
    th, lnl = model.transform_loglike_batch(U)
    assert np.array_equal(th, model.prior_transform_batch(U))
    assert np.allclose(lnl, _single_device_blocks(case, th, 1), rtol=0.0, atol=5e-10)
    c = model.counters()
    assert c["n_points"] == B + B + 1 + 1000
    model.close()


def test_gather_wait_is_bounded_and_the_handle_recovers():
    """A peer that never signals: the fused all-gather returns RVL_EPEER after the timeout instead
    of spinning forever, names the missing rank, and rvl_reset makes the handle usable again."""
    import torch
    from evidence_b200 import synth
    from evidence_b200.rvmodel import DeviceError, RVModel
    case = synth.make_case(2, n_epochs=200)
    rows = 256
    model = RVModel(case.fixedpardict, case.datadict(), case.parnames, device=0)
    model.set_option("gather_timeout_ms", 50)
    theta = case.draw_theta(rows, seed=1)
    # two "ranks" on one GPU: our own gathered vector and a stand-in for a peer that is dead
    mine = torch.zeros(2 * rows + 2, dtype=torch.float64, device="cuda")
    dead = torch.zeros(2 * rows + 2, dtype=torch.float64, device="cuda")
    out = np.empty(2 * rows)
    with pytest.raises(DeviceError, match="EPEER.*0x2"):
        model.log_likelihood_gather_host(theta, out, [mine.data_ptr(), dead.data_ptr()], 0,
                                         2 * rows, 1)
    # our own block did arrive (the kernel ran to its end)
    want = model.log_likelihood_batch(theta)
    assert np.array_equal(out[:rows], want)
    model.reset()
    # "rank 1" signals by hand this time: the call completes
    mine.view(torch.int64)[2 * rows + 1] = 2
    model.log_likelihood_gather_host(theta, out, [mine.data_ptr(), dead.data_ptr()], 0, 2 * rows, 2)
    assert np.array_equal(out[:rows], want)
    model.close()


def test_evidence_ladder_one_run_per_gpu(tmp_path):
    """BASELINE.json configs[3]: the ladder's runs are replicas, one per GPU, no collective; the same
    seeded sampler on the CPU checker (tests/ladder_cpu.py) gives ln Z within the reported uncertainty
    -- in fact the identical chain."""
    import json
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(root, "examples", "evidence_ladder.py"), "--kmax", "1", "--epochs", "120",
           "--nlive", "100", "--true-planets", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    recs = {r["k"]: r for r in (json.loads(l) for l in out.stdout.splitlines()
                                if l.startswith("{") and '"k"' in l)}
    assert sorted(recs) == [0, 1]
    assert {r["device"] for r in recs.values()} == {0, 1}  # one run per GPU
    from ladder_cpu import cpu_ladder
    for c in cpu_ladder(1, epochs=120, nlive=100, true_planets=1):
        r = recs[c["k"]]
        assert np.isfinite(r["logz"]) and abs(r["logz"] - c["cpu_logz"]) <= max(r["logzerr"], c["cpu_logzerr"]), (r, c)
        assert abs(r["logz"] - c["cpu_logz"]) < 0.05 and r["ncall"] == c["cpu_ncall"]  # the same chain
