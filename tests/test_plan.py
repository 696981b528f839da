"""CPU tests of the launch plan (rvl_plan_describe): the graded work list of the likelihood kernel.

The kernel decodes a queue index into (point, sub-slice of the block's resident epoch range); the
same decode is replayed here in Python over every index of every queue, and the items must cover
each (point, 32-epoch chunk) exactly once, with non-overlapping partial-sum slots and arrival
counters.  No device is needed: the planner is plain host code behind the C-ABI.
"""
import ctypes

import numpy as np
import pytest

from evidence_b200 import _abi

SMEM = 232448  # sharedMemPerBlockOptin of a B200
SMS = 148


def describe(Ctot, B, ncol=4, wstride=22, U=2, W=28, sm=SMS, smem=SMEM, sched=1, slices=0,
             items_per_warp=4, min_chunks=8, phase_items=200, max_split=8):
    lib = _abi.load()
    arr = (ctypes.c_int32 * 13)(Ctot, ncol, wstride, U, W, sm, smem, sched, slices,
                                items_per_warp, min_chunks, phase_items, max_split)
    out = (ctypes.c_int64 * 64)()
    rc = lib.rvl_plan_describe(arr, B, out, 64)
    assert rc == 0
    o = list(out)
    plan = dict(Sm=o[0], cpm=o[1], grid=o[2], nph=o[3], nitems=o[4], ptS0=o[5], n_split=o[6],
                partial=o[7])
    plan["phases"] = [dict(idx0=o[8 + 5 * i], S=o[9 + 5 * i], cps=o[10 + 5 * i], pt0=o[11 + 5 * i],
                           part0=o[12 + 5 * i]) for i in range(plan["nph"])]
    return plan


def replay(plan, Ctot, B, U):
    """What the kernel does with the plan; returns per-point chunk coverage counts."""
    Sm, cpm = plan["Sm"], plan["cpm"]
    assert plan["grid"] % Sm == 0 and plan["grid"] >= Sm and plan["grid"] <= max(SMS, Sm)
    assert cpm % U == 0 and (Sm - 1) * cpm < Ctot <= Sm * cpm
    cover = np.zeros((B, Ctot), dtype=np.int32)
    arrive = np.zeros(max(1, plan["n_split"]), dtype=np.int64)
    slots = set()
    finished = np.zeros(B, dtype=np.int32)
    ph = plan["phases"]
    assert ph[0]["idx0"] == 0 and ph[0]["pt0"] == 0
    for sl in range(Sm):
        c0 = sl * cpm
        nch = min(cpm, Ctot - c0)
        assert nch >= 1
        for idx in range(plan["nitems"]):
            k = 0
            while k + 1 < len(ph) and idx >= ph[k + 1]["idx0"]:
                k += 1
            S, cps = ph[k]["S"], ph[k]["cps"]
            j = idx - ph[k]["idx0"]
            jp, ss = divmod(j, S)
            pt = ph[k]["pt0"] + jp
            assert 0 <= pt < B
            lo = ss * cps
            hi = min(nch, lo + cps)
            assert lo % U == 0
            if hi > lo:
                cover[pt, c0 + lo:c0 + hi] += 1
            Stot = Sm * S
            if Stot == 1:
                finished[pt] += 1
            else:
                a = pt - plan["ptS0"]
                assert 0 <= a < plan["n_split"]
                slot = ph[k]["part0"] + (jp * Stot + sl * S + ss) * 2
                assert slot not in slots and slot + 2 <= plan["partial"]
                slots.add(slot)
                arrive[a] += 1
                if arrive[a] == Stot:
                    finished[pt] += 1
                    arrive[a] = 0
    return cover, finished, arrive


CASES = [
    # (Ctot, B, U, kwargs)
    (32, 4096, 2, {}),               # bench: config 2, ndraw = 4096
    (32, 4096, 1, {"W": 32}),
    (32, 1, 2, {}),
    (32, 7, 2, {}),
    (32, 300, 2, {}),
    (7, 100, 2, {}),                 # N = 200: odd chunk count
    (1, 50, 2, {}),
    (2, 50, 1, {"W": 32}),
    (157, 3000, 2, {"ncol": 3}),     # config 3: N = 5000, resident in one SM
    (313, 2000, 2, {"ncol": 3}),     # config 5: N = 10000 needs two resident ranges
    (625, 500, 2, {"ncol": 3}),      # N = 20000: three ranges
    (32, 4096, 2, {"sched": 0}),     # uniform slice count
    (32, 4096, 2, {"slices": 5}),
    (313, 700, 2, {"ncol": 3, "slices": 7}),
    (32, 4096, 2, {"phase_items": 100, "max_split": 16}),
    (32, 20000, 2, {"phase_items": 50, "max_split": 4}),
    (157, 60000, 4, {"ncol": 3, "W": 16}),   # the bench build: 4 epochs per lane, 512 threads
    (313, 5000, 4, {"ncol": 3, "W": 16}),    # stress shape, two resident ranges, U = 4
    (32, 4096, 3, {"W": 20}),
    (7, 100, 4, {"W": 16}),
]


@pytest.mark.parametrize("Ctot,B,U,kw", CASES)
def test_items_cover_every_point_and_chunk_once(built_lib, Ctot, B, U, kw):
    plan = describe(Ctot, B, U=U, **kw)
    cover, finished, arrive = replay(plan, Ctot, B, U)
    assert (cover == 1).all()
    assert (finished == 1).all()      # exactly one item writes lnL
    assert (arrive == 0).all()        # arrival counters are left re-armed
    S_seq = [p["S"] for p in plan["phases"]]
    assert S_seq == sorted(S_seq)     # coarse -> fine


def test_bench_plan_is_graded(built_lib):
    plan = describe(32, 4096)
    assert plan["Sm"] == 1 and plan["cpm"] == 32 and plan["grid"] == SMS
    S_seq = [p["S"] for p in plan["phases"]]
    assert S_seq == [2, 4, 8]
    # every split phase holds about two items per warp of the chip
    last = plan["phases"][-1]
    n_last = plan["nitems"] - last["idx0"]
    assert 1.5 * SMS * 28 <= n_last <= 2.5 * SMS * 28
    big = describe(32, 65536)
    assert [p["S"] for p in big["phases"]] == [1, 2, 4, 8] and big["ptS0"] > 58000


def test_large_batches_split_only_the_tail(built_lib):
    plan = describe(157, 1_000_000, ncol=3)
    assert plan["Sm"] == 1
    assert plan["n_split"] < 20000 and plan["ptS0"] > 980000
    assert plan["partial"] < 2_000_000


def test_rejects_bad_input(built_lib):
    lib = _abi.load()
    arr = (ctypes.c_int32 * 13)(0, 4, 22, 2, 28, SMS, SMEM, 1, 0, 4, 8, 100, 16)
    out = (ctypes.c_int64 * 64)()
    assert lib.rvl_plan_describe(arr, 10, out, 64) != 0
    assert lib.rvl_plan_describe(None, 10, out, 64) != 0


def test_random_plans_cover_exactly_once(built_lib):
    """Fuzz: 150 random (epochs, batch, kernel build, option) combinations."""
    rng = np.random.default_rng(2024)
    for _ in range(150):
        Ctot = int(rng.choice([1, 2, 3, 7, 8, 31, 32, 33, 64, 157, 313, 400]))
        B = int(rng.choice([1, 2, 5, 17, 64, 100, 257, 1000]))
        if Ctot * B > 120_000:
            B = max(1, 120_000 // Ctot)
        U = int(rng.choice([1, 2, 3, 4]))
        kw = dict(W=int(rng.choice([16, 24, 28, 32])), ncol=int(rng.choice([3, 4, 6])),
                  wstride=int(rng.choice([12, 22, 40, 90])), sm=int(rng.choice([8, 132, 148])),
                  sched=int(rng.choice([0, 1])), slices=int(rng.choice([0, 0, 0, 2, 3, 9])),
                  items_per_warp=int(rng.choice([1, 4, 8])), min_chunks=int(rng.choice([1, 4, 8])),
                  phase_items=int(rng.choice([10, 100, 200, 400])),
                  max_split=int(rng.choice([1, 2, 8, 16, 64])))
        plan = describe(Ctot, B, U=U, **kw)
        Sm, cpm = plan["Sm"], plan["cpm"]
        assert plan["grid"] % Sm == 0 and cpm % U == 0 and (Sm - 1) * cpm < Ctot <= Sm * cpm, (Ctot, B, U, kw)
        cover = np.zeros((B, Ctot), dtype=np.int32)
        fin = np.zeros(B, dtype=np.int32)
        arrive = np.zeros(max(1, plan["n_split"]), dtype=np.int64)
        ph = plan["phases"]
        for sl in range(Sm):
            c0 = sl * cpm
            nch = min(cpm, Ctot - c0)
            for idx in range(plan["nitems"]):
                k = 0
                while k + 1 < len(ph) and idx >= ph[k + 1]["idx0"]:
                    k += 1
                jp, ss = divmod(idx - ph[k]["idx0"], ph[k]["S"])
                pt = ph[k]["pt0"] + jp
                lo = ss * ph[k]["cps"]
                hi = min(nch, lo + ph[k]["cps"])
                if hi > lo:
                    cover[pt, c0 + lo:c0 + hi] += 1
                Stot = Sm * ph[k]["S"]
                if Stot == 1:
                    fin[pt] += 1
                else:
                    a = pt - plan["ptS0"]
                    assert 0 <= a < plan["n_split"]
                    arrive[a] += 1
                    if arrive[a] == Stot:
                        fin[pt] += 1
                        arrive[a] = 0
        assert (cover == 1).all() and (fin == 1).all() and (arrive == 0).all(), (Ctot, B, U, kw)
